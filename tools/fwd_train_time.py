"""Times the surrogate-training step at batch B (default 65536) with the per-section profile (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic
from pigan_b200.fwd_trainer import ForwardTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(42)
F = ForwardModel(4, 250, 8)
tr = ForwardTrainer(F, "cuda", max_batch=B)
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device="cuda")
for _ in range(3): tr.step(pn, sp, mn, 1e-3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): tr.step(pn, sp, mn, 1e-3)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 2 * 4132352 * B
print(f"surrogate train step B={B}: {ms:.3f} ms -> {B / ms * 1e3:.3e} samples/s, {flop / ms / 1e9:.0f} TFLOP/s; losses {tr.losses.tolist()}")
tr.engine.profile_begin()
for _ in range(n): tr.step(pn, sp, mn, 1e-3)
for k, (c, t) in tr.engine.profile_end().items():
    print(f"  {k:24s} {c:5d} {t / n * 1e3:8.1f} us/step")
