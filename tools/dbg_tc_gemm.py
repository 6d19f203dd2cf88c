"""GPU check of the tcgen05 3xTF32 GEMM for the three operand-major combinations (run via gpurun)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "distributed-drl_b200"))
import torch
from ddrl_b200 import _native as N

lib = N.lib()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)


def run(M, N_, K, a_mn, b_mn, splits=1):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev, generator=g)
    B = torch.randn((K, N_) if b_mn else (N_, K), device=dev, generator=g)
    Cc = torch.full((splits, M, N_), float("nan"), device=dev)
    N.check(lib.ddrl_debug_tc_gemm(0, C.c_void_p(A.data_ptr()), A.shape[0], A.shape[1], int(a_mn), C.c_void_p(B.data_ptr()),
                                   B.shape[0], B.shape[1], int(b_mn), C.c_void_p(Cc.data_ptr()), M, N_, K, splits, None))
    torch.cuda.synchronize()
    opA = (A.t() if a_mn else A).double()
    opB = (B if b_mn else B.t()).double()
    want = opA @ opB
    got = Cc.double().sum(0)
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"M={M} N={N_} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} splits={splits}: rel err {err:.3e}", flush=True)
    return err


for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 1), (1, 0)):
    for (M, N_, K) in ((128, 128, 32), (128, 128, 256), (256, 256, 64), (300, 200, 100), (37, 17, 29)):
        run(M, N_, K, a_mn, b_mn)
run(257, 256, 1024, 1, 1, splits=4)
