"""Physics-metrics kernel on N synthetic spectra (BASELINE config 3): spectra/s and achieved HBM GB/s."""
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
def run():
    native.check(native.lib.pigan_physics_metrics(spec.data_ptr(), n, 250, freq.data_ptr(), None, 0.0, idx.data_ptr(),
                                                  out.data_ptr(), native.current_stream()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"physics n={n}: {ms:.3f} ms -> {n / ms * 1e3:.3e} spectra/s, {n * 1016 / ms / 1e6:.0f} GB/s")
assert torch.equal(idx.long(), spec.argmin(dim=1)), "peak index differs from argmin"
