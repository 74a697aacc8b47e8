"""Two-CTA GEMM (pigan_debug_linear2) vs torch and vs the one-CTA kernel: numerics and time."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from pigan_b200 import native_test as native
_out = {}
def run(fn, a, b, bias, m, n, k, leaky, fresh=False):
    if fresh or (m, n) not in _out:
        _out[(m, n)] = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float16)
    out = _out[(m, n)]
    native.check(fn(a.data_ptr(), None, b.data_ptr(), native.ptr(bias), out.data_ptr(), None, m, n, k, leaky, native.current_stream()))
    return out
for (m, n, k) in [(256, 256, 64), (1000, 512, 256), (65536, 512, 256), (65536, 256, 512), (65536, 512, 512), (65536, 1024, 512), (65536, 512, 1024)]:
    torch.manual_seed(m + n + k)
    a = (torch.randn(m, k, device="cuda") * 0.5).half()
    b = (torch.randn(n, k, device="cuda") * 0.1).half()
    bias = torch.randn((n + 255) // 256 * 256, device="cuda")
    ref = torch.nn.functional.leaky_relu(a.float() @ b.float().t() + bias[:n], 0.2)
    res = {}
    for name, fn in (("1cta", native.lib.pigan_debug_linear), ("2cta", native.lib.pigan_debug_linear2)):
        out = run(fn, a, b, bias, m, n, k, 1, fresh=True)
        torch.cuda.synchronize()
        err = float((out.float() - ref).norm() / ref.norm())
        for _ in range(3): run(fn, a, b, bias, m, n, k, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run(fn, a, b, bias, m, n, k, 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[name] = (err, ms)
    print(f"M={m} N={n} K={k}: " + "  ".join(f"{nm}: rel {e:.2e} {ms*1e3:.1f} us {2*m*n*k/ms/1e9:.0f} TF/s" for nm, (e, ms) in res.items()))
