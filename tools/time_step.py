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
for _ in range(3): tr.step(sp, pr, mn, 2e-4, 2e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): tr.step(sp, pr, mn, 2e-4, 2e-4)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"train step B={B}: {ms:.3f} ms  -> {B / ms * 1e3:.3e} samples/s ; losses {tr.losses.tolist()[:3]}")
if len(sys.argv) > 2:
    tr.engine.profile_begin()
    for _ in range(n): tr.step(sp, pr, mn, 2e-4, 2e-4)
    for k, (c, t) in tr.engine.profile_end().items():
        print(f"  {k:24s} {c:5d} {t / n * 1e3:8.1f} us/step")
G.eval()
st = flat.net_state(G, "generator")
for _ in range(3): tr.engine.score_candidates(st.params.tensor(), st.bn.tensor(), spectra=sp, want_params=False)
torch.cuda.synchronize()
e0.record()
for _ in range(n): tr.engine.score_candidates(st.params.tensor(), st.bn.tensor(), spectra=sp, want_params=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"score B={B}: {ms:.3f} ms -> {B / ms * 1e3:.3e} cand/s")
tr.engine.profile_begin()
for _ in range(n): tr.engine.score_candidates(st.params.tensor(), st.bn.tensor(), spectra=sp, want_params=False)
for k, (c, t) in tr.engine.profile_end().items():
    print(f"  {k:24s} {c:5d} {t / n * 1e3:8.1f} us/call")
