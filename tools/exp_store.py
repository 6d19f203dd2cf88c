import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import torch, ctypes as C
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import ReplayBuffer, _native
dev = torch.device("cuda")
for name, D, A, cap in [("C2", 24, 4, 2_000_000), ("C3", 376, 17, 1_000_000)]:
    rb = ReplayBuffer(D, A, cap, seed=1)
    n = 1_000_000
    src = [torch.randn(n, D, device=dev), torch.rand(n, A, device=dev), torch.randn(n, device=dev), torch.randn(n, D, device=dev), torch.zeros(n, device=dev)]
    lib = _native.lib(); s = torch.cuda.current_stream()
    def go():
        _native.check(lib.ddrl_rb_store_batch(rb._h, *[C.c_void_p(t.data_ptr()) for t in src], n, 0, C.c_void_p(s.cuda_stream)))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): go()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    row = 4 * (2 * D + A + 2)
    print(name, os.environ.get("DDRL_ROW_ALIGN"), f"{t*1e6:.1f} us  {2*row*n/t/1e9:.0f} GB/s algorithmic", flush=True)
    del rb
