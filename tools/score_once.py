"""ncu target: two scoring calls (generator eval + surrogate + fused reconstruction error) at batch 65 536."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic, flat
from pigan_b200.trainer import NativeTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
torch.manual_seed(42)
G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8); F.eval()
tr = NativeTrainer(G, D, F, dev, max_batch=B)
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device=dev)
G.eval()
st = flat.net_state(G, "generator")
for _ in range(2):
    r = tr.engine.score_candidates(st.params.tensor(), st.bn.tensor(), spectra=sp, want_params=False)
torch.cuda.synchronize()
print("ok", float(r["recon_error"].mean()))
