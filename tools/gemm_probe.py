import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pigan_b200 import native_test as native
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for M in (65536, 148*128):
  for variant in (10, 11, 0):
    for n, k in [(512, 256), (512, 1024), (256, 4096)]:
        a = torch.randn(M, k, device="cuda").half(); b = torch.randn(n, k, device="cuda").half()
        c = torch.empty(M, n, device="cuda")
        st = native.current_stream()
        f = lambda: native.check(native.lib.pigan_debug_gemm_tn(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, n, k, variant, st))
        ms = timeit(f)
        print(dict(M=M, variant=variant, N=n, K=k, ms=round(ms,4), tflops=round(2*M*n*k/ms/1e9,1)), flush=True)
