"""ncu target: three PI-GAN train steps at the BASELINE config-5 widths (generator 2048-2048-2048-4, discriminator
2052-2048-2048-1, surrogate 4-2048x5-2056, batch 65 536 by default) on one GPU.  19 tcgen05 GEMM launches per step:
`ncu --set full --clock-control none -k regex:gemm_tc -s 38 -c 19 -o rep python tools/wide_once.py` captures the third
step's; `python tools/ncu_summary.py rep.ncu-rep` turns the report into the table kept under profiles/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))

import torch

from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from core.models.generator import Generator
from pigan_b200.trainer import NativeTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
torch.manual_seed(11)
G, D = Generator(2048, 4, hidden=(2048, 2048)), Discriminator(2048, 4, hidden=(2048, 2048))
F = ForwardModel(4, 2048, 8, hidden=(2048,) * 5).eval()
tr = NativeTrainer(G, D, F, dev, max_batch=B)
g = torch.Generator(device=dev)
g.manual_seed(1)
x = -3.0 * torch.rand(B, 2048, device=dev, generator=g)
p = 2.2 + 0.6 * torch.rand(B, 4, device=dev, generator=g)
m = torch.rand(B, 8, device=dev, generator=g)
for _ in range(3):
    losses = tr.step(x, p, m, 2e-4, 2e-4)
torch.cuda.synchronize()
print("losses", [round(v, 5) for v in losses.tolist()])
