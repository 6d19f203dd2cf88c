"""GPU check of the tcgen05 3xTF32 GEMM alone, printed (the asserted version is tests/test_tc_gemm_gpu.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200"), os.path.join(ROOT, "tests")]
from ddrl_b200 import _native as N
from test_tc_gemm_gpu import run

for bn in (64, 128):
    for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 1), (1, 0)):
        for (M, N_, K) in ((128, 128, 32), (128, 128, 256), (256, 256, 64), (300, 200, 100), (37, 17, 29)):
            print(f"bn={bn} M={M} N={N_} K={K} a_mn={a_mn} b_mn={b_mn}: rel err {run(N, M, N_, K, a_mn, b_mn, bn):.3e}", flush=True)
    print(f"bn={bn} split-K: {run(N, 257, 256, 1024, 1, 1, bn, splits=4):.3e}", flush=True)
