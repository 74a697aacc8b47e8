"""Times the tcgen05 GEMM core (test-hook epilogues) on the layer shapes of the hot path."""
import sys, os, json
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
res = []
for variant, n, k in [(10, 512, 512), (12, 512, 512), (12, 512, 256)]:
    a = torch.randn(M, k, device="cuda").half(); b = torch.randn(n, k, device="cuda").half()
    c = torch.empty(M, n, device="cuda")
    st = native.current_stream()
    f = lambda: native.check(native.lib.pigan_debug_gemm_tn(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, n, k, variant, st))
    ms = timeit(f)
    ref = timeit(lambda: torch.matmul(a, b.t()))
    res.append(dict(kind="tn", variant=variant, M=M, N=n, K=k, ms=ms, tflops=2*M*n*k/ms/1e9, cublas_ms=ref, cublas_tflops=2*M*n*k/ref/1e9))
    print(res[-1], flush=True)
for n, k, bias, leaky, rs in [(512, 256, False, False, False), (512, 256, True, True, False), (256, 512, True, False, False), (512, 512, True, False, False), (1024, 512, True, False, True), (512, 1024, True, False, True), (256, 512, True, False, True)]:
    a = torch.randn(M, k, device="cuda").half(); b = torch.randn(n, k, device="cuda").half()
    bias_t = torch.randn((n + 255) // 256 * 256, device="cuda") if bias else None; out = torch.empty(M, n, device="cuda", dtype=torch.float16)
    rowst = torch.zeros(M, (n + 255) // 256, 2, device="cuda") if rs else None
    st = native.current_stream()
    f = lambda: native.check(native.lib.pigan_debug_linear(a.data_ptr(), None, b.data_ptr(), native.ptr(bias_t), out.data_ptr(), native.ptr(rowst), M, n, k, int(leaky), st))
    ms = timeit(f)
    if k <= 256:   # production keeps the weights resident for K <= 256: also time the streamed-operand kernel
        native.lib.pigan_debug_force_streamed(1)
        ms_streamed = timeit(f)
        native.lib.pigan_debug_force_streamed(0)
        print(f"   streamed-operand kernel for the same shape: {ms_streamed:.4f} ms (resident: {ms:.4f} ms)", flush=True)
    ref = timeit(lambda: torch.nn.functional.linear(a, b))
    res.append(dict(kind="linear", bias=bias, leaky=leaky, rowstats=rs, M=M, N=n, K=k, ms=round(ms,4), tflops=round(2*M*n*k/ms/1e9,1), hbm_gbs=round((M*k*2+M*n*2)/ms/1e6), cublas_ms=round(ref,4), cublas_tflops=round(2*M*n*k/ref/1e9,1)))
    print(res[-1], flush=True)
for kd, m, n, sp in [(65536, 512, 256, 37), (131072, 256, 512, 37), (65536, 256, 512, 37), (65536, 512, 256, 74)]:
    a = torch.randn(kd, m, device="cuda").half(); b = torch.randn(kd, n, device="cuda").half()
    c = torch.zeros(m, n, device="cuda")
    st = native.current_stream()
    f = lambda: native.check(native.lib.pigan_debug_gemm_nt(a.data_ptr(), b.data_ptr(), None, c.data_ptr(), kd, m, n, sp, 0, 0, 0, -1, None, st))
    ms = timeit(f)
    ref = timeit(lambda: torch.matmul(a.t(), b))
    res.append(dict(kind="nt", Kd=kd, M=m, N=n, splits=sp, ms=ms, tflops=2*kd*m*n/ms/1e9, cublas_ms=ref, cublas_tflops=2*kd*m*n/ref/1e9))
    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_perf.json", "w"), indent=1)
