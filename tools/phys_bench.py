"""Physics-metrics kernel on N synthetic spectra (BASELINE config 3): spectra/s and achieved HBM GB/s, staged
(bulk asynchronous copies) against the register-prefetch kernel (PIGAN_PHYS_BULK=0), forward and backward."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from pigan_b200 import native, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
base, _, _, _ = synthetic.make_batch(1 << 16, 250, seed=3, device="cuda")
spec = base.repeat(max(1, n >> 16), 1)[:n].contiguous()
freq = synthetic.frequencies(250, device="cuda")
idx = torch.empty(n, device="cuda", dtype=torch.int32)
out = torch.empty(n, 4, device="cuda", dtype=torch.float32)
gm = torch.ones(n, 4, device="cuda"); gs = torch.empty_like(spec)
def fwd():
    native.check(native.lib.pigan_physics_metrics(spec.data_ptr(), n, 250, freq.data_ptr(), None, 0.0, idx.data_ptr(),
                                                  out.data_ptr(), native.current_stream()))
def bwd():
    native.check(native.lib.pigan_physics_metrics_backward(spec.data_ptr(), n, 250, freq.data_ptr(), None, 0.0,
                                                           gm.data_ptr(), gs.data_ptr(), None, None, native.current_stream()))
for name, fn, byt in (("forward", fwd, 1016), ("backward", bwd, 2016)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"physics {name} n={n} bulk={os.environ.get('PIGAN_PHYS_BULK', '1')}: {ms:.3f} ms -> {n / ms * 1e3:.3e} spectra/s, {n * byt / ms / 1e6:.0f} GB/s")
assert torch.equal(idx.long(), spec.argmin(dim=1)), "peak index differs from argmin"
