"""bench.py — the learner-side hot path on N GPUs of one node (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C1|C2|C3] [--only-primary]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one iteration of the reference's learner hot loop (algos/sac1/sac1.py:146-148):
    batch = replay_buffer.sample_batch(B);  agent.train(batch)
on synthetic transitions of the named shape.  The headline (`value`, `ms_per_step`, `e2e`, `roofline`) is config C2
(BASELINE.json configs[1]: obs 24, act 4, replay 1e6 rows per GPU, batch 1024, 256x256 MLPs); at N = 1 the same JSON
line also carries `configs: {C1, C2, C3, C4_naive, C4_dedup}` so that every named shape is measured by the same run.
Per rank: one replay shard + one learner; N > 1 adds the gradient exchange (fused into the optimiser kernel over
NVLink peer memory; DDRL_DP=nccl: NCCL all-reduce).

Printed JSON line (rank 0):
  value / unit        whole-job transitions/s consumed by the learners (N * B * K / t), everything resident in HBM,
                      timed with CUDA events between barriers, max over ranks
  e2e                 the same metric through the reference-facing API with HOST buffers: per step
                      sample_batch(B) returned as host numpy arrays -> train(host batch) -> store_batch(B new rows from
                      pinned host memory) -> read the losses back
  roofline            the dominant kernel, measured live with CUDA events: one forward stage of the SAC1 update
                      (fwd_fused_tc: 4 passes x [first + second layer] in one tcgen05 launch; wide inputs: the
                      second-layer gemm_grouped_tc launch), algorithmic FLOPs against the measured bf16 tensor peak;
                      traffic from profiles/ncu_traffic.json (tools/ncu_traffic.py regenerates it from ncu captures)
  roofline_step       the whole SAC1 update by the FLOP model of SURVEY §8d
  roofline_replay     the sample_batch gather kernel alone at the largest launch of the sweep
  configs             per named shape: value, ms_per_step, e2e, gather / store GB/s and fraction of the HBM peak against
                      rows per launch, the numpy reference ring on one host thread beside it
  c5                  the same loop with 256 producers' rows stored every step and the parameter-server broadcast
                      every 300 steps (BASELINE config 5 flavour)
  cpu_baseline        the reference's CPU path on this box's cores: numpy ReplayBuffer (the reference's own class when
                      oracle/_ref was materialised, else its restatement) + the torch-CPU float32 SAC1 step (TensorFlow
                      1.x cannot be installed): all cores (`value`), one thread (`threads1`), replay alone (`replay_only`)
--impl reference times that CPU path as the reference arm.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "distributed-drl_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: obs_dim, act_dim, replay rows per GPU, batch, hidden, act_scale
    "C1": dict(D=8, A=2, replay=1_000_000, B=256, hidden=(256, 256), act_scale=1.0,
               desc="SAC1 LunarLanderContinuous-v2 shapes"),
    "C2": dict(D=24, A=4, replay=1_000_000, B=1024, hidden=(256, 256), act_scale=1.0,
               desc="SAC1 BipedalWalker-shaped synthetic transitions"),
    "C3": dict(D=376, A=17, replay=10_000_000, B=4096, hidden=(256, 256), act_scale=0.4,
               desc="SAC1 Humanoid-shaped synthetic transitions"),
}
# C4: Atari-shaped uint8 84x84x4 frame replay, 1e6 transitions over 8 GPUs = 125 000 per shard, batch 512
C4 = dict(frame=(84, 84), stack=4, shard=125_000, naive_rows=30_000, B=512)


def flops_per_update(D, A, h1, h2, B):
    """5*L_pi + 10*L_q."""
    l_pi = 2 * B * (D * h1 + h1 * h2 + 2 * h2 * A)
    l_q = 2 * B * ((D + A) * h1 + h1 * h2 + h2)
    return 5 * l_pi + 10 * l_q


def ncu_traffic():
    """dram bytes read + written per launch of each config's dominant kernel, from the committed table
    (profiles/ncu_traffic.json, regenerated from ncu --set full captures by tools/ncu_traffic.py)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tf=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm_gbs=6650.0, bf16_tf=1400.0, source="B200_PROFILING.md fallback")


def make_opt(cfg, seed=0):
    space = SimpleNamespace(high=np.full(cfg["A"], cfg["act_scale"], np.float32), shape=(cfg["A"],))
    return SimpleNamespace(obs_dim=cfg["D"], act_dim=cfg["A"], ac_kwargs=dict(hidden_sizes=cfg["hidden"], action_space=space),
                           alpha=0.2, gamma=0.99, lr=1e-3, polyak=0.995, seed=seed, batch_size=cfg["B"])


def workload(name, cfg, per_gpu=True):
    return (f"{name}: {cfg['desc']} (obs {cfg['D']}, act {cfg['A']}, batch {cfg['B']}" + (" per GPU" if per_gpu else "") +
            f", {cfg['hidden'][0]}x{cfg['hidden'][1]} MLP)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.dev), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), sm_mhz_min=min(sm) if sm else None)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path on host cores
# ------------------------------------------------------------------------------------------------
def reference_ring_class():
    """The reference's own numpy ReplayBuffer (algos/sac1/sac1.py:28-63), materialised under oracle/_ref by
    oracle/materialize_ref.py in the build container — else the restatement oracle/replay_oracle.py."""
    from oracle import materialize_ref
    p = materialize_ref.path("ref_replay_sac1.py")
    if p:
        ns = {}
        with open(p) as f:
            exec(compile(f.read(), p, "exec"), ns)
        sha = (materialize_ref.manifest() or {}).get("ref_replay_sac1.py", {}).get("sha256", "?")
        return ns["ReplayBuffer"], "reference", f"the reference's own class, algos/sac1/sac1.py (sha256 {sha[:12]})"
    from oracle.replay_oracle import ReplayRingOracle
    return (lambda obs_dim, act_dim, size: ReplayRingOracle(obs_dim, act_dim, size)), "port", "oracle/replay_oracle.py restatement"


def cpu_ring(cfg, rows, seed=1002):
    cls, kind, what = reference_ring_class()
    D, A = cfg["D"], cfg["A"]
    g = np.random.Generator(np.random.PCG64(seed))
    ring = cls(obs_dim=D, act_dim=A, size=rows)
    ring.obs1_buf[:] = g.standard_normal((rows, D), dtype=np.float32)
    ring.obs2_buf[:] = g.standard_normal((rows, D), dtype=np.float32)
    ring.acts_buf[:] = g.uniform(-1, 1, (rows, A)).astype(np.float32)
    ring.rews_buf[:] = g.standard_normal(rows, dtype=np.float32)
    ring.done_buf[:] = (g.random(rows) < 0.01).astype(np.float32)
    ring.size, ring.ptr = rows, 0
    return ring, kind, what, g


def cpu_replay_only(cfg, budget_s=1.5, ring_rows=None):
    """numpy ring alone, ONE host thread (what a Ray actor process gives the reference): sample_batch(B) and store()."""
    rows = int(ring_rows or min(cfg["replay"], 1_000_000))
    ring, kind, what, g = cpu_ring(cfg, rows)
    B, D, A = cfg["B"], cfg["D"], cfg["A"]
    np.random.seed(7)
    for _ in range(3):
        ring.sample_batch(B)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        ring.sample_batch(B)
        n += 1
    ts = (time.perf_counter() - t0) / n
    o, a, o2 = g.standard_normal(D), g.uniform(-1, 1, A), g.standard_normal(D)
    m, t0 = 20000, time.perf_counter()
    for _ in range(m):
        ring.store(o, a, 0.5, o2, False)
    tst = (time.perf_counter() - t0) / m
    row_bytes = 4 * (2 * D + A + 2)
    return dict(kind=kind, what=what, cores=1, ring_rows=rows, sample_batch_us=ts * 1e6, sample_transitions_per_s=B / ts,
                sample_gbs=2 * row_bytes * B / ts / 1e9, store_us=tst * 1e6, store_transitions_per_s=1.0 / tst,
                sample="%d sample_batch(%d) calls, %d store() calls" % (n, B, m))


def cpu_port_run(cfg, steps, warmup, threads, ring_rows=None, seed=1002):
    """reference numpy ring sample_batch + torch-CPU float32 SAC1 step (oracle/), `steps` timed iterations.
    Returns (seconds per step, description, kind of the replay part)."""
    import torch
    from oracle.sac1_oracle import SAC1Oracle, init_params
    torch.set_num_threads(threads)
    D, A, B = cfg["D"], cfg["A"], cfg["B"]
    rows = int(ring_rows or min(cfg["replay"], 1_000_000))
    ring, kind, what, g = cpu_ring(cfg, rows, seed)
    learner = SAC1Oracle(D, A, hidden=cfg["hidden"], params=init_params(D, A, cfg["hidden"], seed), dtype=torch.float32,
                         act_scale=cfg["act_scale"])
    np.random.seed(seed)

    def one():
        batch = ring.sample_batch(B)
        noise = g.standard_normal((3, B, A), dtype=np.float32)
        learner.step(batch, noise)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, (f"{steps} steps of sample_batch({B}) [{what}] + float32 SAC1 update [torch-CPU port: TensorFlow 1.x is not "
                f"installable], ring {rows} rows, {threads} torch threads"), kind


def cpu_baseline_block(cfg, B):
    cores = os.cpu_count() or 1
    probe, _, _ = cpu_port_run(cfg, 2, 1, cores)
    n = int(max(3, min(200, 10.0 / max(probe, 1e-4))))
    sec_cpu, sample, kind = cpu_port_run(cfg, n, 1, cores)
    n1 = int(max(3, min(100, 6.0 / max(probe, 1e-4))))
    sec_1, sample_1, _ = cpu_port_run(cfg, n1, 1, 1)
    return dict(value=B / sec_cpu, unit="transitions/s", cores=cores, kind="port", sample=sample, ms_per_step=sec_cpu * 1e3,
                replay_part=kind,
                threads1=dict(value=B / sec_1, unit="transitions/s", cores=1, ms_per_step=sec_1 * 1e3, sample=sample_1,
                              note="the reference pins TensorFlow to one intra/inter-op thread (actor_learner.py:110-111)"),
                replay_only=cpu_replay_only(cfg))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 60))
    sec, sample, kind = cpu_port_run(cfg, steps, max(1, min(args.warmup, 5)), cores)
    value = cfg["B"] / sec
    line = dict(metric="learner-path transitions/s (replay sample_batch -> SAC1 update)", value=value, unit="transitions/s",
                impl="reference", n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 5), ms_per_step=sec * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                updates_per_s=1.0 / sec,
                config=dict(workload=workload(args.config, cfg), replay_rows_per_gpu=min(cfg["replay"], 1_000_000)),
                cpu_baseline=dict(value=value, unit="transitions/s", cores=cores, kind="port", sample=sample, replay_part=kind,
                                  replay_only=cpu_replay_only(cfg)),
                e2e=dict(value=value, unit="transitions/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="reference arm = the reference's CPU path: its own numpy ReplayBuffer class (when oracle/_ref was "
                     "materialised in the build container) + the torch-CPU float32 restatement of the SAC1 step; the "
                     "reference's TensorFlow 1.x / Ray stack is not installable here")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if self.rank == 0:
            __graft_entry__.build()
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            dist.barrier()
        from ddrl_b200 import _native
        self.N = _native
        self.lib = _native.lib()
        self.dev = torch.device("cuda", self.local)
        self.peaks = measured_peaks()
        self.traffic = ncu_traffic()

    # -- plumbing ------------------------------------------------------------------------------------
    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, also_wait=()):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = self.N.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        for side in also_wait:                  # work queued on other streams belongs to the timed region
            torch.cuda.current_stream(self.local).wait_stream(side)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = self.N.launch_count() - before
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / 1e3, launches

    def stream(self):
        return self.torch.cuda.current_stream(self.local)

    def fill(self, rb, rows, D, A):
        torch = self.torch
        gen = torch.Generator(device=self.dev).manual_seed(1002 + self.rank)
        chunk = 250_000
        for lo in range(0, rows, chunk):
            n = min(chunk, rows - lo)
            rb.store_batch(torch.randn((n, D), device=self.dev, generator=gen), torch.rand((n, A), device=self.dev, generator=gen) * 2 - 1,
                           torch.randn(n, device=self.dev, generator=gen), torch.randn((n, D), device=self.dev, generator=gen),
                           (torch.rand(n, device=self.dev, generator=gen) < 0.01).float())

    def row_stride_bytes(self, rb):
        rf = C.c_int()
        self.N.check(self.lib.ddrl_rb_layout(rb.native_handle, None, None, C.byref(rf), None))
        return 4 * int(rf.value)

    # -- the gather / store kernels alone, against rows per launch --------------------------------------
    def replay_sweep(self, rb, cfg, sweep=(1, 64, 2048)):
        torch = self.torch
        D, A, B, rows = cfg["D"], cfg["A"], cfg["B"], cfg["replay"]
        row_bytes = 4 * (2 * D + A + 2)
        nbs = sorted({nb for nb in sweep if nb * B * row_bytes <= 1.6e9} | {1})
        nmax = max(nbs) * B
        f32 = dict(dtype=torch.float32, device=self.dev)
        o = [torch.empty((nmax, D), **f32), torch.empty((nmax, D), **f32), torch.empty((nmax, A), **f32),
             torch.empty(nmax, **f32), torch.empty(nmax, **f32)]
        src = [torch.randn((nmax, D), **f32), torch.rand((nmax, A), **f32), torch.randn(nmax, **f32),
               torch.randn((nmax, D), **f32), torch.zeros(nmax, **f32)]
        s = self.stream()
        sp = C.c_void_p(s.cuda_stream)
        gather, store = [], []
        for nb in nbs:
            # `reps` launches back to back between ONE pair of events: the host enqueues ahead of the GPU, so no launch
            # latency of the host sits inside the timed region
            reps = max(10, min(300, 3000 // nb))
            for i in range(3 + reps):
                if i == 3:
                    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
                self.N.check(self.lib.ddrl_rb_sample(rb.native_handle, B, nb, None, 77, i, self.rank,
                                                     *[C.c_void_p(t.data_ptr()) for t in o], None, sp))
            e1 = torch.cuda.Event(enable_timing=True); e1.record(s); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 1e3 / reps
            gbs = 2 * row_bytes * nb * B / t / 1e9
            gather.append(dict(batches_per_launch=nb, rows_per_launch=nb * B, us_per_launch=t * 1e6, transitions_per_s=nb * B / t,
                               gbs=gbs, frac=gbs / self.peaks["hbm_gbs"]))
            m = min(nb * B, rows)
            a = [t_[:m] for t_ in src]
            for i in range(3 + reps):
                if i == 3:
                    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
                self.N.check(self.lib.ddrl_rb_store_batch(rb.native_handle, *[C.c_void_p(t_.data_ptr()) for t_ in a], m, 0, sp))
            e1 = torch.cuda.Event(enable_timing=True); e1.record(s); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 1e3 / reps
            gbs = 2 * row_bytes * m / t / 1e9
            store.append(dict(rows_per_launch=m, us_per_launch=t * 1e6, transitions_per_s=m / t, gbs=gbs,
                              frac=gbs / self.peaks["hbm_gbs"]))
        del o, src
        return dict(bytes_per_transition=2 * row_bytes, row_bytes=row_bytes, gather=gather, store=store,
                    note="algorithmic bytes = 2 x row_bytes per transition (index bytes excluded) over the CUDA-event time "
                         "of the kernel launches alone, against the measured HBM copy peak")

    # -- one SAC1 shape: device-resident step, e2e step, replay kernels, dominant kernel -----------------
    def measure(self, name, steps, warmup, primary):
        from ddrl_b200 import Learner, ReplayBuffer
        torch = self.torch
        cfg = CONFIGS[name]
        D, A, B, hidden, rows = cfg["D"], cfg["A"], cfg["B"], cfg["hidden"], cfg["replay"]
        row_bytes = 4 * (2 * D + A + 2)
        rb = ReplayBuffer(D, A, rows, device=self.local, seed=1000 + 2, rng_stream=self.rank)
        self.fill(rb, rows, D, A)
        stride = self.row_stride_bytes(rb)
        learner = Learner(make_opt(cfg, seed=7), "learner", device=self.local)   # same seed -> same initial weights on every rank
        dp_mode = "single"
        if self.world > 1:
            dp_mode = "nccl"
            if os.environ.get("DDRL_DP", "fused") != "nccl" and learner.connect_peers():
                dp_mode = "peer-fused-nvls" if getattr(learner, "nvls", False) else "peer-fused"
        torch.cuda.synchronize()

        # (1) device-resident hot loop: sample_batch(B) + train(batch) as one native call (the step's first kernel gathers
        #     the batch from the ring; bit-identical to train(rb.sample_batch(B, device=True)), tests/test_sac_gpu.py)
        clocks = ClockSampler(self.local)
        if self.rank == 0 and primary:
            clocks.start()
            time.sleep(0.05)
        sec, launches = self.timed(lambda: learner.train_from_buffer(rb, B), steps, warmup)
        clk = clocks.stop() if (self.rank == 0 and primary) else None
        out = dict(workload=workload(name, cfg), value=self.world * B * steps / sec, unit="transitions/s",
                   ms_per_step=sec / steps * 1e3, updates_per_s=steps / sec, steps=steps, gpu_launches=int(launches),
                   replay_rows_per_gpu=rows, row_bytes=row_bytes, row_stride_bytes=stride, replay_bytes_per_gpu=rows * stride)

        # (2) end to end through the reference-facing API with host buffers
        g = np.random.Generator(np.random.PCG64(5 + self.rank))
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        new = [pin(g.standard_normal((B, D), dtype=np.float32)), pin(g.uniform(-1, 1, (B, A)).astype(np.float32)),
               pin(g.standard_normal(B, dtype=np.float32)), pin(g.standard_normal((B, D), dtype=np.float32)),
               pin((g.random(B) < 0.01).astype(np.float32))]
        sink = []

        # Three shapes of the reference's learner loop, all with host numpy batches crossing PCIe both ways, this step's B new
        # transitions stored from host arrays, and the losses fetched every step:
        #   model   example/model.py:92-101 `Model.train(replay_buffer, args)`: sample into host memory + feed from host memory as
        #           one method (Learner.train_via_host: two queued native calls, nothing waits in between)      <- `e2e`
        #   cache   algos/sac1/sac1.py:136-151 `batch = cache.q1.get(); agent.train(batch)` with the prefetching Cache
        #   blocking  `batch = replay_buffer.sample_batch(B); agent.train(batch)` with no prefetch
        from ddrl_b200 import Cache

        # the rollout side stores on ITS OWN stream, as the reference's workers are their own processes: the buffer orders
        # the store kernel against the learner's samples with events, so the H2D copy of the new rows overlaps the update
        rollout_stream = torch.cuda.Stream(device=self.dev)

        def store_new():
            with torch.cuda.stream(rollout_stream):
                rb.store_batch(*new)                              # H2D: B new transitions from the rollout side

        def step_e2e():
            res = learner.train_via_host(rb, B)                   # D2H batch -> host numpy, H2D batch, update
            store_new()
            sink.append(res["scalars"].cpu())                     # D2H: the fetched losses

        cache = Cache(rb, B, depth=2)

        def step_e2e_cache():
            batch = cache.q1.get()                                # host numpy arrays, prefetched like the reference's Cache
            res = learner.train(batch)
            store_new()
            sink.append(res["scalars"].cpu())

        prev = [None]

        def step_e2e_pipelined():                                 # the cache loop with the losses of step k fetched after step
            batch = cache.q1.get()                                # k + 1 has been queued (one read-back per step, one step late)
            res = learner.train(batch)
            store_new()
            if prev[0] is not None:
                sink.append(prev[0]())                        # waits for step k - 1 only
            prev[0] = learner.losses_async()

        def step_e2e_blocking():
            batch = rb.sample_batch(B)
            res = learner.train(batch)
            store_new()
            sink.append(res["scalars"].cpu())

        e2e_steps = max(5, min(steps, 200 if primary else 60))
        e2e_warm = max(3, min(warmup, 10))
        sec_e2e, _ = self.timed(step_e2e, e2e_steps, e2e_warm, also_wait=(rollout_stream,))
        sec_blk, _ = self.timed(step_e2e_blocking, e2e_steps, e2e_warm, also_wait=(rollout_stream,))
        cache.start()
        sec_cache, _ = self.timed(step_e2e_cache, e2e_steps, e2e_warm, also_wait=(rollout_stream,))
        sec_pipe, _ = self.timed(step_e2e_pipelined, e2e_steps, e2e_warm, also_wait=(rollout_stream,))
        cache.end()
        sub = lambda sec, path: dict(value=self.world * B * e2e_steps / sec, ms_per_step=sec / e2e_steps * 1e3, path=path)
        out["e2e"] = dict(value=self.world * B * e2e_steps / sec_e2e, unit="transitions/s", h2d_bytes_per_step=2 * B * row_bytes,
                          d2h_bytes_per_step=B * row_bytes + 16, steps=e2e_steps, ms_per_step=sec_e2e / e2e_steps * 1e3,
                          path="Learner.train_via_host(replay_buffer, B) [example/model.py:92-101 Model.train(replay_buffer, args): "
                               "sample_batch into host numpy arrays, update fed from those host arrays] -> store_batch(host, B new "
                               "rows, on the rollout side's own CUDA stream) -> losses.cpu(); batches of <= 2 MB cross PCIe inside the gather kernel / the "
                               "step's first kernel (the pinned block's device mapping) instead of as separate DMA copies",
                          cache_loop=sub(sec_cache, "Cache(replay_buffer).q1.get() -> numpy batch -> Learner.train(numpy) -> "
                                                    "store_batch(host) -> losses.cpu() (algos/sac1/sac1.py:136-151; 2 samples in flight)"),
                          pipelined=sub(sec_pipe, "the cache loop with each step's losses read back after the NEXT step has been "
                                                  "queued (still one D2H loss read per step): the GPU never waits for the host"),
                          blocking=sub(sec_blk, "sample_batch() -> numpy -> Learner.train(numpy) -> store_batch(host) -> losses.cpu()"))
        extra = {}
        if primary:
            # (2b) config C5 flavour: 256 vectorised rollout producers store CONCURRENTLY with the learner — a producer
            #      thread on its own CUDA stream pushes one host row per producer (pinned host arrays -> H2D -> store kernel)
            #      for every learner step while the learner's stream runs sample -> update; weights go to the actors'
            #      parameter-server replica every 300 learner steps (sac1.py:149) by ONE broadcast
            import threading
            from ddrl_b200.dist import DistributedParameterServer
            producers = 256
            rows_per_rank = max(1, producers // self.world)
            prod = [pin(g.standard_normal((rows_per_rank, D), dtype=np.float32)), pin(g.uniform(-1, 1, (rows_per_rank, A)).astype(np.float32)),
                    pin(g.standard_normal(rows_per_rank, dtype=np.float32)), pin(g.standard_normal((rows_per_rank, D), dtype=np.float32)),
                    pin(np.zeros(rows_per_rank, np.float32))]
            keys, values = learner.get_weights()
            ps = DistributedParameterServer(keys, values, src=0, device=self.dev)
            st = dict(i=0, stored=0, err=None)
            permits, done_evt, quit_evt = threading.Semaphore(0), threading.Event(), threading.Event()
            pstream = torch.cuda.Stream(device=self.dev)

            def producer_loop():
                try:
                    with torch.cuda.stream(pstream):
                        while True:
                            permits.acquire()
                            if quit_evt.is_set():
                                return
                            rb.store_batch(*prod)                 # H2D + store kernel on the producers' stream
                            st["stored"] += rows_per_rank
                except BaseException as e:                        # noqa: BLE001
                    st["err"] = e
                finally:
                    done_evt.set()

            th = threading.Thread(target=producer_loop, daemon=True)
            th.start()

            def step_c5():
                permits.release()                                 # this step's 256 / N transitions arrive from the rollout side
                learner.train_from_buffer(rb, B)
                st["i"] += 1
                if st["i"] % 300 == 0:
                    ps.push_flat(learner.get_flat_weights())      # device-to-device, then one NCCL broadcast of 0.88 MB
                    ps.sync()

            def drain():                                          # producers have issued everything they were asked for
                while st["stored"] < st["i"] * rows_per_rank and st["err"] is None:
                    time.sleep(0)
                pstream.synchronize()

            c5_steps = max(300, min(steps, 600))
            c5_warm = max(3, min(warmup, 10))
            for _ in range(c5_warm):
                step_c5()
            drain()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(c5_steps):
                step_c5()
            drain()                                               # the timed region ends when learner AND producers are done
            e1.record()
            self.barrier()
            wall = time.perf_counter() - t0
            sec_c5 = e0.elapsed_time(e1) / 1e3
            if self.world > 1:
                t = torch.tensor([sec_c5], device=self.dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                sec_c5 = float(t.item())
            quit_evt.set(); permits.release(); th.join(5)
            if st["err"] is not None:
                raise st["err"]
            extra["c5"] = dict(value=self.world * B * c5_steps / sec_c5, unit="transitions/s", ms_per_step=sec_c5 / c5_steps * 1e3,
                               steps=c5_steps, producers=producers, stored_rows_per_step=rows_per_rank * self.world,
                               stored_transitions_per_s=self.world * rows_per_rank * c5_steps / sec_c5,
                               h2d_bytes_per_step=rows_per_rank * row_bytes, wall_ms_per_step=wall / c5_steps * 1e3,
                               ps_broadcast_every=300,
                               note="config 5 flavour of the same loop, end to end on the producer side: a producer thread per rank "
                                    "stores its share of 256 producers' transitions from pinned HOST arrays on its own CUDA "
                                    "stream (reservation serialised by the buffer, kernels ordered by events) while the learner "
                                    "stream samples and updates; every 300 steps the flat weights are broadcast to the "
                                    "parameter-server replicas")
        # (3) the gather / store kernels alone against rows per launch
        out["replay"] = self.replay_sweep(rb, cfg, sweep=(1, 8, 64, 512, 2048) if primary else (1, 64, 2048))
        # (4) the dominant kernel alone, live: one forward stage of the update, CUDA events on the launching stream
        s = self.stream()
        sp = C.c_void_p(s.cuda_stream)
        reps = 100

        def stage_time(r):      # r > 0: back-to-back stream launches; r < 0: the same launches as the nodes of one graph
            self.N.check(self.lib.ddrl_sac_debug_stage(learner._h, B, 1, r, sp))        # warm-up (r < 0: also builds the graph)
            torch.cuda.synchronize()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record(s)
            self.N.check(self.lib.ddrl_sac_debug_stage(learner._h, B, 1, r, sp))
            t1e.record(s)
            torch.cuda.synchronize()
            return t0e.elapsed_time(t1e) / 1e3 / abs(r)

        # the step runs as ONE CUDA graph, so the kernel is timed as graph nodes; as separate stream launches every kernel is
        # rounded up to the next ~2.05 us completion tick of stream-order launches (tools/probes/tick_probe.cu)
        tc_s, tc_stream_s = stage_time(-reps), stage_time(reps)
        h1, h2 = hidden
        fused = D + A <= 32 and h1 <= 256 and h1 % 32 == 0 and h2 % 4 == 0
        if fused:   # passes a, c (policy, K1 = D) and d, e (Q, K1 = D + A): first + second layer
            flops = 2 * B * (2 * (D * h1 + h1 * h2) + 2 * ((D + A) * h1 + h1 * h2))
            kname = ("fwd_fused_tc: first + second layer of 4 forward passes in one launch (TMA -> tcgen05.mma kind::tf32 x3, "
                     "layer-2 A operand from TMEM -> TMA store)")
        else:       # second-layer launch of the same four passes
            flops = 2 * B * 4 * h1 * h2
            kname = "gemm_grouped_tc: second layer of 4 forward passes in one launch (TMA -> tcgen05.mma kind::tf32 x3 -> TMEM -> TMA store)"
        tr = self.traffic.get(name, {})
        out["roofline"] = dict(kernel=kname, bound="tensor", achieved=flops / tc_s / 1e12, peak=self.peaks["bf16_tf"], unit="TFLOP/s",
                               frac=flops / tc_s / 1e12 / self.peaks["bf16_tf"], traffic=tr.get("dram_bytes"),
                               tensor_pipe_active_pct=tr.get("tensor_pipe_active_pct"), traffic_source=tr.get("source"),
                               us_per_launch=tc_s * 1e6, us_per_stream_launch=tc_stream_s * 1e6, timed_as="nodes of one CUDA graph "
                               "(how the step runs them); us_per_stream_launch = the same kernel as back-to-back stream launches",
                               flops_per_launch=flops, tensor_pipe_flops_per_launch=3 * flops,
                               peak_source=self.peaks["source"])
        fl = flops_per_update(D, A, h1, h2, B)
        tf = fl * (steps / sec) / 1e12
        out["roofline_step"] = dict(bound="tensor", achieved=tf, peak=self.peaks["bf16_tf"], unit="TFLOP/s", frac=tf / self.peaks["bf16_tf"],
                                    traffic=None, flops_per_update=fl, tf32x3_tensor_flops_per_update=3 * fl)
        extra["clocks"], extra["dp_mode"] = clk, dp_mode
        del learner, rb
        torch.cuda.empty_cache()
        return out, extra

    # -- C4: Atari-shaped uint8 frame replay, one 125 000-transition shard per GPU -----------------------
    def measure_frames(self):
        from ddrl_b200.frames import FrameReplayBuffer
        torch = self.torch
        fb = C4["frame"][0] * C4["frame"][1]
        res = {}

        def timeit(fn, iters=30, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e-3

        sp = C.c_void_p(self.stream().cuda_stream)
        ptr = lambda t: C.c_void_p(t.data_ptr())

        def both(api_call, native_call, nb, alg):
            """host API (fresh output tensors per call, Python wrapper) and the kernel alone (native calls, preallocated
            outputs, back to back between one pair of events) for a batch of nb transitions of alg bytes each"""
            t_api, t_k = timeit(api_call), timeit(native_call, iters=100)
            mk = lambda t: dict(us=t * 1e6, transitions_per_s=nb / t, gbs=alg * nb / t / 1e9, frac=alg * nb / t / 1e9 / self.peaks["hbm_gbs"])
            return dict(batch=nb, kernel=mk(t_k), host_api=mk(t_api), **mk(t_k))

        rb = FrameReplayBuffer(C4["frame"], C4["stack"], C4["shard"], mode="dedup", seed=1, device=self.local)
        t_store = []
        z = torch.zeros(25_000, device=self.dev)
        for lo in range(0, C4["shard"], 25_000):
            fr = torch.randint(0, 256, (25_000, fb), dtype=torch.uint8, device=self.dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rb.store_frames(fr, z, z, z); e1.record(); torch.cuda.synchronize()
            t_store.append(e0.elapsed_time(e1) * 1e-3)
        alg_d = 5 * fb + 12 + 2 * 4 * fb + 12               # read 5 frames + scalars, write two stacks + scalars
        gather = []
        for nb in (C4["B"], 8192):
            o1 = torch.empty((nb, 4 * fb), dtype=torch.uint8, device=self.dev); o2 = torch.empty_like(o1)
            sc = torch.empty((3, nb), dtype=torch.float32, device=self.dev)
            cnt = [0]

            def native():
                cnt[0] += 1
                self.N.check(self.lib.ddrl_fb_sample_stack(self.local, ptr(rb.frames), fb, C4["stack"], rb.max_size, rb.size, 0,
                                                           ptr(rb.act), ptr(rb.rew), ptr(rb.done), nb, None, 1, cnt[0], 0,
                                                           ptr(o1), ptr(o2), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), None, sp))
            gather.append(both(lambda: rb.sample_batch(nb), native, nb, alg_d))
            del o1, o2, sc
        st = min(t_store[1:]) if len(t_store) > 1 else t_store[0]
        res["C4_dedup"] = dict(workload="C4: Atari-shaped uint8 84x84x4 frame replay, frame-deduplicated ring (one frame per transition, "
                                        "stacks rebuilt from 5 consecutive frames), 125 000-transition shard (1e6 over 8 GPUs), batch 512",
                               value=gather[0]["host_api"]["transitions_per_s"], unit="transitions/s (sample_batch(512) calls back to back, host API)",
                               bytes_per_transition=alg_d, gather=gather,
                               store=dict(rows_per_launch=25_000, us=st * 1e6, gbs=2 * 25_000 * (fb + 12) / st / 1e9,
                                          frac=2 * 25_000 * (fb + 12) / st / 1e9 / self.peaks["hbm_gbs"]))
        del rb
        torch.cuda.empty_cache()
        rbn = FrameReplayBuffer(C4["frame"], C4["stack"], C4["naive_rows"], mode="naive", seed=1, device=self.local)
        for lo in range(0, C4["naive_rows"], 5_000):
            o = torch.randint(0, 256, (5_000, 4) + C4["frame"], dtype=torch.uint8, device=self.dev)
            z5 = torch.zeros(5_000, device=self.dev)
            rbn.store_batch(o, z5, z5, o, z5)
        alg_n = 2 * (2 * 4 * fb + 12)
        gather = []
        for nb in (C4["B"], 8192):
            o1 = torch.empty((nb, fb), dtype=torch.float32, device=self.dev); o2 = torch.empty_like(o1)
            sc = torch.empty((3, nb), dtype=torch.float32, device=self.dev)
            cnt = [0]

            def native_n():
                cnt[0] += 1
                self.N.check(self.lib.ddrl_rb_sample(rbn._rb.native_handle, nb, 1, None, 1, cnt[0], 0, ptr(o1), ptr(o2), ptr(sc[0]),
                                                     ptr(sc[1]), ptr(sc[2]), None, sp))
            gather.append(both(lambda: rbn.sample_batch(nb), native_n, nb, alg_n))
            del o1, o2, sc
        res["C4_naive"] = dict(workload="C4: the same frames with obs1 + obs2 stored per transition (56 460-byte rows), 30 000 rows "
                                        "(1.7 GB, larger than L2), batch 512",
                               value=gather[0]["host_api"]["transitions_per_s"], unit="transitions/s (sample_batch(512) calls back to back, host API)",
                               bytes_per_transition=alg_n, gather=gather)
        del rbn
        torch.cuda.empty_cache()
        return res


def measure_qlearn(b):
    """Row N4: the DDQN / SQN learner steps at the reference's default shape (hidden [400, 300], batch 256; obs 115 = the
    dqn family's trading observation, 3 actions), device-resident, with the float32 torch-CPU restatement beside them."""
    from types import SimpleNamespace
    from ddrl_b200 import DQNLearner, SQNLearner
    from oracle.qlearn_oracle import DDQNOracle, SQNOracle, init_q_params, make_q_batch
    torch = b.torch
    D, nA, hidden, B = 115, 3, (400, 300), 256
    opt = SimpleNamespace(obs_dim=D, act_dim=nA, hidden_size=list(hidden), gamma=0.99, lr=1e-3, polyak=0.995, seed=0, batch_size=B, alpha=0.1)
    batch = {k: torch.from_numpy(v).to(b.dev) for k, v in make_q_batch(D, nA, B, seed=1).items()}
    out = {}
    for name, cls, ocls, nn in (("ddqn", DQNLearner, DDQNOracle, 1), ("sqn", SQNLearner, SQNOracle, 2)):
        L = cls(opt, "learner", device=b.local)
        sec, launches = b.timed(lambda: L.train(batch), 300, 10)
        ora = ocls(init_q_params(D, nA, hidden, nn, seed=2), alpha=0.1, dtype=torch.float32)
        hb = make_q_batch(D, nA, B, seed=1)
        torch.set_num_threads(os.cpu_count() or 1)
        ora.step(hb)
        t0 = time.perf_counter()
        for _ in range(20):
            ora.step(hb)
        cpu = (time.perf_counter() - t0) / 20
        out[name] = dict(us_per_step=sec / 300 * 1e6, transitions_per_s=B * 300 / sec, gpu_launches_per_step=launches / 300,
                         cpu_port_us_per_step=cpu * 1e6, cpu_cores=os.cpu_count())
        del L
    return dict(workload="N4: DDQN (algos/dqn) and SQN (algos/sqn) learner steps, obs 115, 3 actions, hidden 400x300, batch 256, fp32 FFMA",
                **out)


def measure_frames_sharded(b):
    """C4 as configured: 1e6 Atari-shaped transitions sharded over the N GPUs of the node (frame-deduplicated ring, one
    shard per rank, local sampling — no data-path collective), batch 512 per learner; whole-job transitions/s, max over ranks."""
    from ddrl_b200.dist import ShardedFrameReplayBuffer
    torch = b.torch
    fbytes = C4["frame"][0] * C4["frame"][1]
    rb = ShardedFrameReplayBuffer(C4["frame"], C4["stack"], 1_000_000, mode="dedup", device=b.local, seed=1)
    shard = rb.map.cap
    z = torch.zeros(25_000, device=b.dev)
    for lo in range(0, shard, 25_000):
        n = min(25_000, shard - lo)
        rb.store_frames(torch.randint(0, 256, (n, fbytes), dtype=torch.uint8, device=b.dev), z[:n], z[:n], z[:n])
    out = {}
    for nb in (C4["B"], 8192):
        sec, launches = b.timed(lambda: rb.sample_batch(nb), 50, 5)
        alg = nb * (5 * fbytes + 12 + 2 * 4 * fbytes + 12)
        out[str(nb)] = dict(batch_per_gpu=nb, value=b.world * nb * 50 / sec, unit="transitions/s", us_per_call=sec / 50 * 1e6,
                            gbs_per_gpu=alg * 50 / sec / 1e9, frac=alg * 50 / sec / 1e9 / b.peaks["hbm_gbs"], gpu_launches=int(launches))
    del rb
    torch.cuda.empty_cache()
    return dict(workload="C4: Atari-shaped uint8 84x84x4 frame replay, 1e6 transitions sharded over %d GPU(s) (%d per shard), "
                         "frame-deduplicated ring, sample_batch through the host API back to back" % (b.world, shard),
                shard_transitions=shard, bytes_per_transition=5 * fbytes + 12 + 2 * 4 * fbytes + 12, batches=out, scaling="weak")


def run_ours(args):
    b = Bench(args)
    torch, dist = b.torch, b.dist
    primary_name = args.config
    prim, extra = b.measure(primary_name, args.steps, args.warmup, primary=True)
    configs = {primary_name: prim}
    if b.world == 1 and not args.only_primary:
        for name in CONFIGS:
            if name == primary_name:
                continue
            steps = max(50, min(args.steps, 400 if name == "C1" else 150))
            configs[name], _ = b.measure(name, steps, max(3, min(args.warmup, 20)), primary=False)
            configs[name]["cpu_replay_1thread"] = cpu_replay_only(CONFIGS[name], budget_s=1.0)
        configs.update(b.measure_frames())
        configs["N4_qlearn"] = measure_qlearn(b)
    c4_sharded = measure_frames_sharded(b) if not args.only_primary else None
    if b.world > 1:
        dist.barrier()
    if b.rank != 0:
        if b.world > 1:
            dist.destroy_process_group()
        return
    cfg = CONFIGS[primary_name]
    world = b.world
    rep = prim["replay"]["gather"][-1]
    sto = prim["replay"]["store"][-1]
    line = dict(
        metric="learner-path transitions/s (replay sample_batch -> SAC1 update)", value=prim["value"], unit="transitions/s",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=prim["ms_per_step"], higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        updates_per_s=prim["updates_per_s"], global_batch=world * cfg["B"],
        config=dict(workload=workload(primary_name, cfg),
                    replay_rows_per_gpu=prim["replay_rows_per_gpu"], replay_bytes_per_gpu=prim["replay_bytes_per_gpu"],
                    row_bytes=prim["row_bytes"], row_stride_bytes=prim["row_stride_bytes"],
                    parallelism=(f"dp{world}: replay sharded per GPU, gradient all-reduce "
                                 + ("fused into the optimiser kernel: multimem.ld_reduce over an NVLS multicast mapping (the NVSwitch adds)" if extra["dp_mode"] == "peer-fused-nvls"
                                    else "fused into the optimiser kernel over NVLink peer memory (CUDA IPC)" if extra["dp_mode"] == "peer-fused"
                                    else "by NCCL")) if world > 1 else "single GPU",
                    l2="replay ring larger than L2, rows drawn at random; weights (3.5 MB) are L2-resident by design",
                    noise="Philox on device", index_source="Philox on device"),
        clocks=extra["clocks"],
        e2e=prim["e2e"], c5=extra.get("c5"),
        gpu_launches=prim["gpu_launches"],
        roofline=dict(prim["roofline"],
                      note="algorithmic FLOPs (one product per multiply-add) over the CUDA-event launch time, against the measured "
                           "bf16 tensor peak; fp32-class accuracy (1e-5 bar) costs three tf32 MMAs per product at half the bf16 "
                           "rate, so the tensor pipe executes 6x this figure in bf16-equivalents; at these sizes (128 tiles of "
                           "128x64) a launch is bound by its TMA -> MMA -> epilogue latency chain and by the per-instruction floor "
                           "of the small tcgen05.mma shapes (DESIGN.md §4); traffic = dram bytes read + written per launch "
                           "(operands are L2-resident in the real step)"),
        roofline_step=dict(prim["roofline_step"],
                           kernel="SAC1 update, whole step: prologue + 2 fused forward stages + 2 backward tcgen05 stages + 3 row-wise "
                                  "kernels + FFMA narrow weight gradient + Adam/polyak; side-stream bias / skinny gradients",
                           note="the step is a chain of 10 dependent launches of 5-12 us, i.e. latency-bound; fraction is algorithmic "
                                "FLOPs against the bf16 tensor peak as SURVEY 8(d) defines it"),
        roofline_replay=dict(kernel="rb_gather_* (sample_batch)", bound="hbm", achieved=rep["gbs"], peak=b.peaks["hbm_gbs"],
                             unit="GB/s", frac=rep["frac"], traffic=b.traffic.get(primary_name + "_gather", {}).get("dram_bytes"),
                             bytes_per_transition=prim["replay"]["bytes_per_transition"], transitions_per_launch=rep["rows_per_launch"],
                             transitions_per_s=rep["transitions_per_s"], peak_source=b.peaks["source"],
                             store_gbs=sto["gbs"], store_frac=sto["frac"], store_rows_per_launch=sto["rows_per_launch"]),
        configs=configs if world == 1 else None,
        c4_sharded=c4_sharded,
    )
    # ---- CPU baseline on this box's cores (N = 1 only) -------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(cfg, cfg["B"])
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-primary", action="store_true", help="N = 1: skip the other named configs (C1, C3, C4)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
