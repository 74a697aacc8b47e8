"""Three train steps + one 65 536-candidate search call + the model-validation call + the physics kernels (forward and
backward, 1 M spectra) + data-pipeline and evaluator kernels + two surrogate-training steps
(ncu target; tools/launch_summary.py cuts the list into steps)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic, flat, native
from pigan_b200.trainer import NativeTrainer
B = 65536
dev = torch.device("cuda")
torch.manual_seed(42)
G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8); F.eval()
tr = NativeTrainer(G, D, F, dev, max_batch=B)
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device=dev)
for _ in range(3): tr.step(sp, pr, mn, 2e-4, 2e-4)
torch.cuda.synchronize()
G.eval()
st = flat.net_state(G, "generator")
s, i, p = tr.engine.search(st.params.tensor(), st.bn.tensor(), sp[0], 0.01, 7, 0, B, 256)
val = tr.engine.validate(st.params.tensor(), st.bn.tensor(), sp, torch.randn_like(sp), 0.01)
freq = synthetic.frequencies(250, device=dev)
NP = 1 << 20
big = sp.repeat(NP // B, 1).contiguous()
idx = torch.empty(NP, device=dev, dtype=torch.int32); out = torch.empty(NP, 4, device=dev)
native.check(native.lib.pigan_physics_metrics(big.data_ptr(), NP, 250, freq.data_ptr(), None, 0.0, idx.data_ptr(), out.data_ptr(), native.current_stream()))
gm = torch.ones(NP, 4, device=dev); gs = torch.empty_like(big)
native.check(native.lib.pigan_physics_metrics_backward(big.data_ptr(), NP, 250, freq.data_ptr(), None, 0.0, gm.data_ptr(), gs.data_ptr(), None, None, native.current_stream()))
from pigan_b200 import device_data, evalstats
gen_spec, gen_par = torch.empty(NP, 250, device=dev), torch.empty(NP, 4, device=dev)
device_data.generate_spectra(NP, dev, seed=1, frequency=freq, out=gen_spec, params_out=gen_par)
perm = torch.randperm(NP, device=dev)[:B]
rows = device_data.gather_rows(big, perm)
rm = evalstats.RegressionMetrics(250, dev)
rm.update(big, gen_spec)
mse = rm.compute()["mse"]
del big, gs, gen_spec
torch.cuda.synchronize()
from pigan_b200.fwd_trainer import ForwardTrainer
Ft = ForwardModel(4, 250, 8)
ftr = ForwardTrainer(Ft, dev, max_batch=B, engine=tr.engine)
for _ in range(2): ftr.step(pn, sp, mn, 1e-3)
torch.cuda.synchronize()
print("ok", float(s[0]), ftr.losses.tolist())
