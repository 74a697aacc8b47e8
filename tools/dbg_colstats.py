import sys, os
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic, native
from pigan_b200.trainer import NativeTrainer
print(native.LIB_PATH, os.path.getmtime(native.LIB_PATH))
B = 4096
dev = torch.device("cuda")
torch.manual_seed(42)
G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8); F.eval()
tr = NativeTrainer(G, D, F, dev, max_batch=B)
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device=dev)
tr.engine.profile_begin()
tr.step(sp, pr, mn, 2e-4, 2e-4)
print(tr.engine.profile_end())
print(tr.losses.tolist())
