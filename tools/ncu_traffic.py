"""Regenerate profiles/ncu_traffic.json — the per-config DRAM traffic table bench.py's `roofline.traffic` reads — from
`ncu --set full` captures (run where ncu is installed; no GPU needed):

    python tools/ncu_traffic.py [dir with the .ncu-rep files, default gpurun_out] [round tag, default r02]

Expected captures (tools/prof_r02.sh writes them): <tag>_fwd_<cfg>.ncu-rep = one launch of the dominant kernel of the SAC1
update of that config (fwd_fused_tc for C1 / C2, the second-layer gemm_grouped_tc launch for C3), <tag>_gather_<cfg>.ncu-rep
= the sample_batch gather kernel at the largest launch of bench.py's sweep.  Per capture the table keeps
dram__bytes_read.sum + dram__bytes_write.sum of that ONE launch, its duration under ncu and the tensor-pipe activity."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    r = rows[2]

    def val(name):
        if name not in hdr:
            return None, None
        i = hdr.index(name)
        try:
            return float(r[i].replace(",", "")), units[i]
        except ValueError:
            return None, units[i]
    return r[hdr.index("Kernel Name")], val


def to_bytes(v, unit):
    if v is None:
        return None
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return int(round(v * mult))


def to_us(v, unit):
    if v is None:
        return None
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
    table = {}
    for cfg in ("C1", "C2", "C3"):
        for kind, key in (("fwd", cfg), ("gather", cfg + "_gather"), ("store", cfg + "_store")):
            p = os.path.join(d, f"{tag}_{kind}_{cfg.lower()}.ncu-rep")
            if not os.path.isfile(p):
                continue
            kname, val = raw(p)
            rd, wr = to_bytes(*val("dram__bytes_read.sum")), to_bytes(*val("dram__bytes_write.sum"))
            table[key] = dict(kernel=kname[:100], dram_bytes=(rd or 0) + (wr or 0), dram_bytes_read=rd, dram_bytes_write=wr,
                              us_under_ncu=to_us(*val("gpu__time_duration.sum")),
                              tensor_pipe_active_pct=val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[0],
                              dram_throughput_pct=val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[0],
                              grid=val("launch__grid_size")[0], waves_per_sm=val("launch__waves_per_multiprocessor")[0],
                              source=f"profiles/{tag}_{kind}_{cfg.lower()}_ncu_full_summary.txt (ncu --set full, one launch, cold-cache replay)")
    out = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    with open(out, "w") as f:
        json.dump(table, f, indent=1)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
