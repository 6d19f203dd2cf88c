"""Measured maximum errors of one SAC1 update against the float64 oracle, per tensor group, for both GEMM modes at the
full C1 / C2 / C3 sizes (GPU box only) -> gpurun_out/r02_parity_margins.json (copied to profiles/).  `rel` = max|a - b| /
max|b|, the metric of tests/test_sac_gpu.py; the float32 evaluation of the oracle itself is listed beside the kernels."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import Learner
from test_sac_gpu import build_pair, rel
from oracle.sac1_oracle import SAC1Oracle, conditioned_params, make_batch

out = {}
for name, (D, A, B, scale) in dict(C1=(8, 2, 256, 1.0), C2=(24, 4, 1024, 1.0), C3=(376, 17, 4096, 0.4)).items():
    hidden = (256, 256)
    params = conditioned_params(D, A, hidden, seed=300 + D)
    batch, noise = make_batch(D, A, B, seed=400 + D)
    row = {}
    for gemm in ("tc", "ffma"):
        learner, oracle = build_pair(Learner, D, A, hidden, B, params, gemm=gemm, act_scale=scale)
        want_g = oracle.flat_grads(batch, noise)
        want = oracle.step(batch, noise)
        got = learner.train(batch, noise=noise, split=True, sync_outputs=True)
        sc = got["scalars"].cpu().numpy()
        g = learner.get_flat_weights("grad").cpu().numpy()
        solid = np.abs(want_g) > 1e-4 * np.abs(want_g).max()
        r = dict(losses=[float(abs(sc[i] - float(want[k])) / abs(float(want[k]))) for i, k in enumerate(("pi_loss", "q1_loss", "q2_loss"))],
                 q1=rel(got["q1"].cpu().numpy(), want["q1"]), q2=rel(got["q2"].cpu().numpy(), want["q2"]),
                 logp_pi=rel(got["logp_pi"].cpu().numpy(), want["logp_pi"]), gradient=rel(g, want_g))
        for which in ("main", "target"):
            w, ww = learner.get_flat_weights(which).cpu().numpy(), oracle.flat(which)
            r[f"{which}_weights_where_gradient_not_eps_dominated"] = rel(w[solid], ww[solid])
            r[f"{which}_weights_all"] = rel(w, ww)
        row[gemm] = {k: (float(v) if not isinstance(v, list) else v) for k, v in r.items()}
        print(name, gemm, json.dumps(row[gemm]), flush=True)
    o32 = SAC1Oracle(D, A, hidden=hidden, gamma=0.99, polyak=0.995, lr=1e-3, alpha=0.2, act_scale=scale, params=params, dtype=torch.float32)
    g32 = o32.flat_grads(batch, noise)
    o32.step(batch, noise)
    row["float32_oracle"] = dict(gradient=float(rel(g32, want_g)), main_weights_all=float(rel(o32.flat("main"), oracle.flat("main"))))
    print(name, "float32 oracle", json.dumps(row["float32_oracle"]), flush=True)
    out[name] = row
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(bars="tests/test_sac_gpu.py: losses / q / logp <= 1e-5, gradient <= 1e-5, weights <= 1e-5 where |g| > 1e-4 max|g|, <= 5e-5 elsewhere",
               measured=out), open(os.path.join(ROOT, "gpurun_out", "r02_parity_margins.json"), "w"), indent=1)
