"""Times the on-device synthetic-spectrum generator through the C ABI with a preallocated output (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from pigan_b200 import native
for n in (1 << 20, 1 << 22):
    freq = torch.linspace(0.5, 3.0, 250, device="cuda")
    out = torch.empty(n, 250, device="cuda"); par = torch.empty(n, 4, device="cuda")
    def go(seed):
        native.check(native.lib.pigan_generate_spectra(None, par.data_ptr(), freq.data_ptr(), n, 250, 0.1, seed, 0, 1,
                                                       out.data_ptr(), None, native.current_stream()))
    for i in range(10): go(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): go(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"n={n}: {ms:.3f} ms, {n / ms / 1e6:.3f} G spectra/s, {n * 1016 / ms / 1e6:.0f} GB/s")
