"""Exploratory parity report: native CUDA path vs the CPU oracle (oracle/models.py) on seeded inputs."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pi-gan-thz_b200"))
import torch
from core.models.generator import Generator
from core.models.discriminator import Discriminator
from core.models.forward_model import ForwardModel
from pigan_b200 import synthetic, flat
from pigan_b200.trainer import NativeTrainer, LOSS_KEYS
from pigan_b200 import engine as E
from oracle import models as O

def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

def build(seed=42, randomize_norm=True):
    torch.manual_seed(seed)
    G = Generator(250, 4); D = Discriminator(250, 4); F = ForwardModel(4, 250, 8)
    if randomize_norm:
        g = torch.Generator().manual_seed(seed + 1)
        for m in list(G.modules()) + list(F.modules()):
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.LayerNorm)):
                m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
                m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
    return G, D, F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda")
G, D, F = build()
F.eval()
g_sd = {k: v.clone() for k, v in G.state_dict().items()}
d_sd = {k: v.clone() for k, v in D.state_dict().items()}
f_sd = {k: v.clone() for k, v in F.state_dict().items()}
spec, praw, pnorm, mnorm = synthetic.make_batch(B, 250, seed=7, device="cpu")

G.to(dev); D.to(dev); F.to(dev)
with torch.no_grad():
    # ---- module forwards
    G.eval()
    g_eval = G(spec.to(dev))
    o_eval = O.generator_forward(copy.deepcopy(g_sd), spec, training=False)
    print("G eval        rel", rel(g_eval, o_eval), "maxabs", (g_eval.cpu() - o_eval).abs().max().item())
    G.train()
    g_tr = G(spec.to(dev))
    sd2 = copy.deepcopy(g_sd)
    o_tr = O.generator_forward(sd2, spec, training=True)
    print("G train       rel", rel(g_tr, o_tr), "maxabs", (g_tr.cpu() - o_tr).abs().max().item())
    print("  running_mean1 rel", rel(G.main[1].running_mean, sd2["main.1.running_mean"]), "running_var1", rel(G.main[1].running_var, sd2["main.1.running_var"]),
          "rm2", rel(G.main[4].running_mean, sd2["main.4.running_mean"]), "rv2", rel(G.main[4].running_var, sd2["main.4.running_var"]), "nbt", G.main[1].num_batches_tracked.item())
    d_out = D(spec.to(dev), praw.to(dev))
    o_d = O.discriminator_forward(d_sd, spec, praw)
    print("D forward     rel", rel(d_out, o_d), "maxabs", (d_out.cpu() - o_d).abs().max().item())
    fs, fm = F(pnorm.to(dev))
    os_, om = O.forward_model_forward(f_sd, pnorm)
    print("F spectrum    rel", rel(fs, os_), " metrics rel", rel(fm, om))
    # ---- scoring
    G.load_state_dict(g_sd); G.eval()
    st = flat.net_state(G, "generator")
    eng = E.get_engine(dev, B)
    eng.load_forward_model(flat.net_state(F, "forward_model").params.tensor())
    res = eng.score_candidates(st.params.tensor(), st.bn.tensor(), spectra=spec.to(dev))
    op, ov, oe, oc = O.score_candidates(g_sd, f_sd, spec)
    print("score: params rel", rel(res["params_norm"], op), "err rel", rel(res["recon_error"], oe), "cons rel", rel(res["consistency"], oc),
          "viol mismatches", (res["violations"].cpu().long() != ov).sum().item())
    tgt = spec[0]; noise = torch.randn(B, 250, generator=torch.Generator().manual_seed(3))
    res2 = eng.score_candidates(st.params.tensor(), st.bn.tensor(), target=tgt.to(dev), noise=noise.to(dev), sigma=0.01)
    cand = O.noisy_candidates(tgt[None, :], noise, 0.01)
    op2 = O.generator_forward(copy.deepcopy(g_sd), cand, training=False)
    rec2, _ = O.forward_model_forward(f_sd, op2)
    oe2 = ((tgt[None, :] - rec2) ** 2).mean(1)
    print("noisy score: params rel", rel(res2["params_norm"], op2), " spread-rel", ((res2["params_norm"].cpu() - op2).norm() / (op2 - op2.mean(0)).norm()).item(), "err rel", rel(res2["recon_error"], oe2))

# ---- train steps
G, D, F = build()
F.eval()
g_sd = {k: v.clone() for k, v in G.state_dict().items()}
d_sd = {k: v.clone() for k, v in D.state_dict().items()}
tr = NativeTrainer(G, D, F, dev, max_batch=B)
og, od = O.Adam(O.G_TRAINABLE), O.Adam(O.D_TRAINABLE)
for it in range(2):
    sp, pr, pn, mn = synthetic.make_batch(B, 250, seed=100 + it, device="cpu")
    losses = tr.step(sp.to(dev), pr.to(dev), mn.to(dev), 2e-4, 2e-4).cpu()
    g_before = None
    ol, ex = O.train_step(g_sd, d_sd, f_sd, og, od, (sp, pr, pn, None, mn), 2e-4, 2e-4)
    print(f"--- step {it + 1}")
    for k, v in zip(LOSS_KEYS, losses.tolist()):
        ref = ol[k]
        print(f"  {k:22s} native {v:.7f} oracle {ref:.7f} rel {abs(v - ref) / max(abs(ref), 1e-12):.2e}")
    # grads are clipped in place by the native step; compare direction + norm
    for name, gview in zip(flat.GENERATOR_PARAMS, tr.gs.params.views_like(tr.g_grads)):
        ref = ex["g_grads"][name]; coef = min(1.0, 1.0 / (ex["g_grad_norm"] + 1e-6))
        print(f"  g_grad {name:16s} rel {rel(gview, ref * coef):.2e}")
    for name, gview in zip(flat.DISCRIMINATOR_PARAMS, tr.ds.params.views_like(tr.d_grads)):
        ref = ex["d_grads"][name]; coef = min(1.0, 1.0 / (ex["d_grad_norm"] + 1e-6))
        print(f"  d_grad {name:16s} rel {rel(gview, ref * coef):.2e}")
    for name in flat.GENERATOR_PARAMS:
        print(f"  g_param {name:16s} rel {rel(G.state_dict()[name], g_sd[name]):.2e}", end=";")
    print()
    for name in flat.DISCRIMINATOR_PARAMS:
        print(f"  d_param {name:16s} rel {rel(D.state_dict()[name], d_sd[name]):.2e}", end=";")
    print()
    print("  bn rm1", rel(G.main[1].running_mean, g_sd["main.1.running_mean"]), "rv1", rel(G.main[1].running_var, g_sd["main.1.running_var"]), "nbt", G.main[1].num_batches_tracked.item(), int(g_sd["main.1.num_batches_tracked"]))
    print("  pred_params rel", rel(tr.engine._wrap(0,0,torch.float32) if False else torch.zeros(1), torch.ones(1)) if False else "")
