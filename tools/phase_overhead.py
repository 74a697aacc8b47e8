"""One GPU: the PI-GAN step as ONE C call (pigan_train_step) against the same step driven phase by phase from Python
(pigan_train_step_phase x 7, what the data-parallel schedule does between its exchanges).  The difference is the cost of
splitting the step - ctypes calls, lost fusion of reduce+finalize - without any inter-GPU wait."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic
from pigan_b200.trainer import NativeTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
torch.manual_seed(42)
G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8); F.eval()
tr = NativeTrainer(G, D, F, dev, max_batch=B)
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def whole():
    tr.step(sp, pr, mn, 2e-4, 2e-4)

def phased():
    tr.step_count += 1
    a = tr._args(sp, pr, mn, 2e-4, 2e-4)
    for ph in range(7):
        tr.engine.train_step_phase(a, ph)

for name, fn in (("one C call", whole), ("7 phase calls from Python", phased), ("one C call", whole)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    import time
    n = 30
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name:28s}: {e0.elapsed_time(e1) / n:.4f} ms/step on the device, host issue time {(t1 - t0) / n * 1e3:.3f} ms/step")
