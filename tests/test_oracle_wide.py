"""The oracle at widths other than the reference's (CPU).  tests/test_oracle_golden.py pins oracle/models.py to the
reference's own classes and train_pigan at the reference widths; the widened GPU tests (tests/test_gpu_wide.py) use the
same functions at other widths.  This file closes the gap: at a widened stack the oracle's functional step must equal
the step written the way the reference writes it - torch.nn modules (the drop-in classes' nn.Sequential stacks, i.e.
nn.Linear / nn.BatchNorm1d / nn.LayerNorm exactly as in core/models/*.py), nn.BCELoss / nn.MSELoss, autograd,
clip_grad_norm_ and torch.optim.Adam, in the order of core/train/train_pigan.py:114-187."""
import copy
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-gan-thz_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

S, MT = 192, 8
GH, DH, FH = (512, 256), (256, 512), (256, 512, 256, 512, 256)


def _setup(B=96, seed=4):
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from oracle import fixtures
    from oracle import models as O
    gen = torch.Generator().manual_seed(seed)
    g_sd = O.init_generator(S, 4, GH, gen)
    d_sd = O.init_discriminator(S, 4, DH, gen)
    f_sd = O.init_forward_model(4, S, MT, FH, gen=gen)
    for bi, h in zip((1, 4), GH):
        g_sd[f"main.{bi}.weight"] = 1.0 + (torch.rand(h, generator=gen) - 0.5)
        g_sd[f"main.{bi}.bias"] = 0.4 * (torch.rand(h, generator=gen) - 0.5)
    G, D, F = Generator(S, 4, hidden=GH), Discriminator(S, 4, hidden=DH), ForwardModel(4, S, MT, hidden=FH)
    G.load_state_dict(g_sd); D.load_state_dict(d_sd); F.load_state_dict(f_sd)
    spec, praw, pnorm, mnorm = fixtures.make_batch(B, seed=31, num_points=S)
    return (g_sd, d_sd, f_sd), (G, D, F), (spec, praw, pnorm, mnorm)


def test_oracle_forwards_equal_the_torch_modules_at_widened_widths():
    from oracle import models as O
    (g_sd, d_sd, f_sd), (G, D, F), (spec, praw, pnorm, mnorm) = _setup()
    with torch.no_grad():
        G.eval(); F.eval()
        assert torch.allclose(G.main(spec), O.generator_forward(copy.deepcopy(g_sd), spec, False), atol=1e-6)
        G.train()
        sd = copy.deepcopy(g_sd)
        assert torch.allclose(G.main(spec), O.generator_forward(sd, spec, True), atol=1e-6)
        for k in ("main.1.running_mean", "main.1.running_var", "main.4.running_mean", "main.4.running_var"):
            assert torch.allclose(G.state_dict()[k], sd[k], atol=1e-6), k
        assert torch.allclose(D.main(torch.cat((spec, praw), dim=1)), O.discriminator_forward(d_sd, spec, praw), atol=1e-6)
        out = F.model(pnorm)
        rs, rm = O.forward_model_forward(f_sd, pnorm, S)
        assert torch.allclose(out[:, :S], rs, atol=1e-5) and torch.allclose(out[:, S:], rm, atol=1e-5)


def test_oracle_train_step_equals_the_module_step_at_widened_widths():
    """One D-step + G-step with nn modules / autograd / torch.optim.Adam in train_pigan's order vs oracle.train_step."""
    from core.utils import loss as L
    from oracle import models as O
    (g_sd, d_sd, f_sd), (G, D, F), (spec, praw, pnorm, mnorm) = _setup()
    lr = 2e-4
    B = spec.shape[0]
    # ---- the reference's loop body (train_pigan.py:114-187) on the nn.Sequential stacks
    G.train(); D.train(); F.eval()
    opt_g = torch.optim.Adam(G.parameters(), lr=lr, betas=(0.5, 0.999))
    opt_d = torch.optim.Adam(D.parameters(), lr=lr, betas=(0.5, 0.999))
    bce, mse = nn.BCELoss(), nn.MSELoss()

    def dnet(x, p):
        return D.main(torch.cat((x, p), dim=1))

    def denorm(p):
        return (p + 1.0) / 2.0 * 0.6 + 2.2

    opt_d.zero_grad()
    loss_d_real = bce(dnet(spec, praw), torch.ones(B, 1) * 0.9)
    with torch.no_grad():
        fake = denorm(G.main(spec))
    loss_d = loss_d_real + bce(dnet(spec, fake), torch.zeros(B, 1) + 0.1)
    loss_d.backward()
    d_grads = {n: p.grad.clone() for n, p in D.named_parameters()}
    torch.nn.utils.clip_grad_norm_(D.parameters(), 1.0)
    opt_d.step()
    opt_g.zero_grad()
    pred = G.main(spec)
    loss_adv = bce(dnet(spec, denorm(pred)), torch.ones(B, 1))
    with torch.no_grad():
        out = F.model(pred)
        recon, pm = out[:, :S], out[:, S:]
    l_rec, l_met = mse(recon, spec), mse(pm, mnorm)
    l_mx = L.maxwell_equation_loss(recon, None, pred)
    l_lc = L.lc_model_approx_loss(pm[:, 0:1], pm[:, 1:2], pred)
    l_rng = L.structural_param_range_loss(pred)
    loss_g = (loss_adv + 100.0 * l_rec + 10.0 * l_rec + 1.0 * l_met + 1.0 * l_mx + 1.0 * l_lc + 0.1 * l_rng
              + 0.0 * L.bnn_kl_loss(G))
    loss_g.sum().backward()
    g_grads = {n: p.grad.clone() for n, p in G.named_parameters()}
    torch.nn.utils.clip_grad_norm_(G.parameters(), 1.0)
    opt_g.step()
    # ---- the oracle
    og, od = O.Adam(O.G_TRAINABLE), O.Adam(O.D_TRAINABLE)
    g2, d2 = copy.deepcopy(g_sd), copy.deepcopy(d_sd)
    ref, ex = O.train_step(g2, d2, f_sd, og, od, (spec, praw, pnorm, None, mnorm), lr, lr)

    def close(a, b, tol=2e-5):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)) < tol

    for k, v in (("d_losses", loss_d), ("g_losses", loss_g), ("adv_losses", loss_adv), ("recon_spec_losses", l_rec),
                 ("recon_metrics_losses", l_met), ("maxwell_losses", l_mx), ("lc_losses", l_lc),
                 ("param_range_losses", l_rng)):
        assert abs(float(v.detach().sum()) - ref[k]) <= 2e-5 * abs(ref[k]) + 1e-8, (k, float(v.detach().sum()), ref[k])
    for n in O.D_TRAINABLE:
        assert close(d_grads[n], ex["d_grads"][n]), n
        assert close(D.state_dict()[n], d2[n]), n
    gnorm = float(torch.cat([v.reshape(-1) for v in ex["g_grads"].values()]).norm())
    for n in O.G_TRAINABLE:
        if n in ("main.0.bias", "main.3.bias"):      # feed a BatchNorm: zero up to rounding noise on both sides
            assert float(g_grads[n].norm()) < 1e-4 * gnorm and float(ex["g_grads"][n].norm()) < 1e-4 * gnorm
            continue
        assert close(g_grads[n], ex["g_grads"][n], 2e-4), n
    for k in ("main.1.running_mean", "main.1.running_var", "main.4.running_mean", "main.4.running_var"):
        assert close(G.state_dict()[k], g2[k]), k
    assert int(G.state_dict()["main.1.num_batches_tracked"]) == int(g2["main.1.num_batches_tracked"]) == 2
