"""Pipeline probes of the tcgen05 GEMM mainloop (no stores): streamed vs resident weights, ring depth."""
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

M = 65536
names = {10: "streamed 4 stages, no epilogue", 12: "streamed 3 stages, no epilogue", 24: "streamed 2 stages, no epilogue",
         11: "streamed 4 stages, TMEM loads only", 20: "resident B, 3 A stages, no epilogue",
         21: "resident B, 4 A stages, no epilogue", 22: "resident B, 6 A stages, no epilogue",
         23: "resident B, 6 A stages, TMEM loads only", 25: "resident B, N tile 128, 8 A stages, no epilogue"}
for n, k in [(512, 256), (256, 256), (512, 512)]:
    a = torch.randn(M, k, device="cuda").half(); b = torch.randn(n, k, device="cuda").half()
    c = torch.empty(M, n, device="cuda")
    st = native.current_stream()
    for variant in (10, 12, 24, 11, 20, 21, 22, 23, 25):
        if variant >= 20 and k > 256: continue
        f = lambda: native.check(native.lib.pigan_debug_gemm_tn(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, n, k, variant, st))
        ms = timeit(f)
        print(f"M={M} N={n} K={k}  {names[variant]:48s} {ms*1e3:7.1f} us  {2*M*n*k/ms/1e9:7.0f} TFLOP/s", flush=True)
