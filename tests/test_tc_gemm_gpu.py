"""GPU parity of the tcgen05 3xTF32 grouped GEMM alone (csrc/sac_gemm_tc.cuh, through ddrl_debug_tc_gemm): every
operand-major combination the SAC1 step uses (forward: A K-major / B MN-major; dgrad: K / K; wgrad: MN / MN),
both tile widths (128 x 128 and 128 x 64), ragged M / N / K, split-K — against a float64 matmul.  The products carry
fp32-class accuracy (a_lo.b_hi + a_hi.b_lo + a_hi.b_hi with tf32 hi parts): 2e-6 of max|C| is asserted."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import _native
    return _native


def run(N, M, N_, K, a_mn, b_mn, bn, splits=1, seed=1):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.randn((K, M) if a_mn else (M, K), device=dev, generator=g)
    B = torch.randn((K, N_) if b_mn else (N_, K), device=dev, generator=g)
    Cc = torch.full((splits, M, N_), float("nan"), device=dev)
    N.check(N.lib().ddrl_debug_tc_gemm(0, C.c_void_p(A.data_ptr()), A.shape[0], A.shape[1], int(a_mn), C.c_void_p(B.data_ptr()),
                                       B.shape[0], B.shape[1], int(b_mn), C.c_void_p(Cc.data_ptr()), M, N_, K, splits, bn, None))
    torch.cuda.synchronize()
    want = (A.t() if a_mn else A).double() @ (B if b_mn else B.t()).double()
    got = Cc.double().sum(0)
    return (got - want).abs().max().item() / want.abs().max().item()


@pytest.mark.parametrize("bn", [64, 128])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N_,K", [(128, 128, 32), (128, 128, 256), (256, 256, 64), (300, 200, 100), (37, 17, 29),
                                    (1024, 256, 256)])
def test_gemm_matches_float64(lib, M, N_, K, a_mn, b_mn, bn):
    assert run(lib, M, N_, K, a_mn, b_mn, bn) <= TOL


@pytest.mark.parametrize("bn", [64, 128])
def test_split_k_wgrad_shape(lib, bn):
    # d[W] = act^T . dZ over a 1024-row batch in four 256-row slices (the step's wgrad layout)
    assert run(lib, 257, 256, 1024, 1, 1, bn, splits=4) <= TOL
    assert run(lib, 28, 256, 1024, 1, 1, bn, splits=4) <= TOL
