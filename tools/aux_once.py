"""ncu target for the kernels outside the train step: inverse-design search (Philox candidates, running top-k),
model-validation call, physics metrics forward / backward (1 M spectra), synthetic-spectrum generator, row gather,
evaluator sums."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic, flat, native, engine as E, device_data, evalstats
B = 65536
dev = torch.device("cuda")
torch.manual_seed(42)
G = Generator(250, 4).to(dev).eval(); F = ForwardModel(4, 250, 8).to(dev).eval()
eng = E.Engine(B, dev)
eng.load_forward_model(flat.net_state(F, "forward_model").params.tensor())
st = flat.net_state(G, "generator")
sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=1, device=dev)
s, i, p = eng.search(st.params.tensor(), st.bn.tensor(), sp[0], 0.01, 7, 0, 2 * B, 1024)
val = eng.validate(st.params.tensor(), st.bn.tensor(), sp, torch.randn_like(sp), 0.01)
freq = synthetic.frequencies(250, device=dev)
NP = 1 << 20
big = sp.repeat(NP // B, 1).contiguous()
idx = torch.empty(NP, device=dev, dtype=torch.int32); out = torch.empty(NP, 4, device=dev)
native.check(native.lib.pigan_physics_metrics(big.data_ptr(), NP, 250, freq.data_ptr(), None, 0.0, idx.data_ptr(), out.data_ptr(), native.current_stream()))
gm = torch.ones(NP, 4, device=dev); gs = torch.empty_like(big)
native.check(native.lib.pigan_physics_metrics_backward(big.data_ptr(), NP, 250, freq.data_ptr(), None, 0.0, gm.data_ptr(), gs.data_ptr(), None, None, native.current_stream()))
gen_spec, gen_par = torch.empty(NP, 250, device=dev), torch.empty(NP, 4, device=dev)
device_data.generate_spectra(NP, dev, seed=1, frequency=freq, out=gen_spec, params_out=gen_par)
rows = device_data.gather_rows(big, torch.randperm(NP, device=dev)[:B])
rm = evalstats.RegressionMetrics(250, dev)
rm.update(big, gen_spec)
print("ok", float(s[0]), rm.compute()["mse"], float(val["stability"].mean()))
