"""Times the evaluator regression sums on [n, 250] fp32 arrays (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from pigan_b200 import evalstats
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
y = torch.randn(n, 250, device="cuda"); p = y + 0.25
rm = evalstats.RegressionMetrics(250, "cuda")
for _ in range(2): rm.update(y, p)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): rm.update(y, p)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"n={n}: {ms:.3f} ms, {2 * n * 1000 / ms / 1e6:.0f} GB/s", rm.compute())
