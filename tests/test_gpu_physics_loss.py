"""Physics-metric loss term of the G-step and the surrogate's vector-Jacobian product it rides on
(A19: core/train/unified_trainer.py:240-256, 325 - a loss on F(G(x)) that reaches G through the frozen F).
The reference has no such term (lambda = 0 reproduces it bit for bit); the oracle is torch autograd over the oracle's
own modules plus the pinned `calculate_peak_parameters` restatement (oracle/physics.py)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pi-gan-thz_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)
DEV = "cuda"


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _setup(n, lam):
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from oracle import fixtures
    from pigan_b200.trainer import NativeTrainer
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    G, D, F = Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)
    G.load_state_dict(g_sd); D.load_state_dict(d_sd); F.load_state_dict(f_sd)
    F.eval()
    tr = NativeTrainer(G, D, F, torch.device(DEV), max_batch=n, lambda_physics_metric=lam)
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=55)
    return tr, (g_sd, d_sd, f_sd), (spec, praw, pnorm, mnorm)


@pytest.mark.parametrize("n", [256, 4096])
def test_surrogate_vjp_matches_autograd(n):
    """pigan_forward_model_vjp against torch autograd through the oracle's forward model (fp32) for a random upstream
    gradient: fp16 operands through six layers and five LayerNorm backward passes - 5e-2 in the norm, the bound of the
    A19 test (the reference's own bf16-autocast path is at 9e-2)."""
    from oracle import models as O
    tr, (g_sd, d_sd, f_sd), (spec, praw, pnorm, mnorm) = _setup(n, 0.0)
    gen = torch.Generator().manual_seed(3)
    g_out = torch.randn(n, 258, generator=gen) / n
    dp = tr.engine.forward_model_vjp(tr.fs.params.tensor(), pnorm.to(DEV), g_out.to(DEV))
    tr.engine.load_forward_model(tr.fs.params.tensor())
    p = pnorm.clone().requires_grad_(True)
    s, m = O.forward_model_forward(f_sd, p)
    (torch.cat([s, m], dim=1) * g_out).sum().backward()
    assert rel(dp, p.grad) < 5e-2, rel(dp, p.grad)


def test_lambda_zero_is_the_reference_step_bit_for_bit():
    tr0, _, (spec, praw, pnorm, mnorm) = _setup(512, 0.0)
    a = tr0.step(spec.to(DEV), praw.to(DEV), mnorm.to(DEV), 2e-4, 2e-4).clone()
    g0 = tr0.gs.params.tensor().clone()
    tr1, _, _ = _setup(512, 0.0)
    tr1.lambda_physics_metric = 0.0
    b = tr1.step(spec.to(DEV), praw.to(DEV), mnorm.to(DEV), 2e-4, 2e-4).clone()
    assert torch.equal(a, b) and torch.equal(g0, tr1.gs.params.tensor())


def test_physics_metric_term_reaches_the_generator():
    """One step with lambda > 0, phases by hand: (1) the reported loss equals the formula evaluated by the pinned CPU
    restatement of calculate_peak_parameters on the engine's own reconstruction; (2) the extra generator-output gradient
    equals autograd through the oracle's F of that loss (5e-2, see above); (3) the change of the generator's
    parameter gradients equals autograd through the oracle's G of that extra output gradient."""
    from oracle import models as O
    from oracle import physics as P
    from pigan_b200 import synthetic
    n, lam = 1024, 50.0
    tr, (g_sd, d_sd, f_sd), (spec, praw, pnorm, mnorm) = _setup(n, lam)
    tr0, _, _ = _setup(n, 0.0)
    sg, pg, mg = spec.to(DEV), praw.to(DEV), mnorm.to(DEV)
    grads = {}
    for name, t in (("pm", tr), ("base", tr0)):
        t.step_count += 1
        a = t._args(sg, pg, mg, 2e-4, 2e-4)
        a.flags = 1
        t._pm_spectrum = sg
        for ph in range(6):
            if ph == 3 and t.lambda_physics_metric > 0:
                p_eng = t.engine.generator_output(n).clone()
                rec_eng = t.engine.forward_model_forward(p_eng)[:, :250].clone()
                t._physics_metric_term(a)
            t.engine.train_step_phase(a, ph)
        torch.cuda.synchronize()
        grads[name] = t.g_grads.clone()
    # (1) the loss, from the engine's reconstruction through the CPU restatement
    freq = synthetic.frequencies(250).numpy()
    _, m_rec = P.physics_batch(rec_eng.cpu().numpy(), freq)
    _, m_tgt = P.physics_batch(spec.numpy(), freq)
    ok = np.isfinite(m_rec).all(1) & np.isfinite(m_tgt).all(1)
    w = np.array(tr.physics_metric_weights)
    loss_ref = float((w * np.where(ok[:, None], m_rec - m_tgt, 0.0) ** 2).sum() / n)
    assert ok.sum() > n // 4
    assert abs(float(tr._pm_loss) - loss_ref) <= 1e-3 * loss_ref, (float(tr._pm_loss), loss_ref)
    # (2) d(loss)/dp through the oracle's F: gradient of the loss wrt the reconstruction from the pinned VJP, row by row
    g_rec = np.zeros((n, 250))
    idx = rec_eng.cpu().numpy().argmin(1)
    for r in np.nonzero(ok)[0]:
        gm = 2.0 * w * (m_rec[r] - m_tgt[r]) / n
        _, g_rec[r] = P.peak_parameters_vjp(freq, rec_eng[r].cpu().numpy().astype(np.float64), int(idx[r]), gm)
    p = p_eng.cpu().clone().requires_grad_(True)
    s, _ = O.forward_model_forward(f_sd, p)
    (s * torch.from_numpy(g_rec).float()).sum().backward()
    dp_ref = lam * p.grad
    assert rel(tr._pm_dp, dp_ref) < 5e-2, rel(tr._pm_dp, dp_ref)
    # (3) through G: the difference of the two runs' generator gradients = J_G^T dp_extra (train-mode BatchNorm)
    g = O._leaf(copy.deepcopy(g_sd), O.G_TRAINABLE)
    pn = O.generator_forward(g, spec, training=True)
    (pn * tr._pm_dp.cpu()).sum().backward()
    views_pm = dict(zip(tr.gs.params.names, tr.gs.params.views_like(grads["pm"])))
    views_b = dict(zip(tr0.gs.params.names, tr0.gs.params.views_like(grads["base"])))
    for name in ("main.6.weight", "main.4.weight", "main.3.weight", "main.0.weight"):
        d = views_pm[name] - views_b[name]
        ref = g[name].grad
        assert float(ref.norm()) > 0
        assert rel(d, ref) < 3e-2, (name, rel(d, ref))


def test_forward_model_module_is_differentiable_in_its_input():
    """The drop-in ForwardModel under autograd: a reconstruction loss on F(p) back-propagates into p through the engine's
    VJP, as the reference's UnifiedTrainer does (unified_trainer.py:240-256, 325); the module's weights get no gradient."""
    from core.models.forward_model import ForwardModel
    from oracle import fixtures
    from oracle import models as O
    _, _, f_sd = fixtures.make_weights(42)
    F = ForwardModel(4, 250, 8)
    F.load_state_dict(f_sd)
    F = F.to(DEV).eval()
    n = 512
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=8)
    p = pnorm.to(DEV).requires_grad_(True)
    s, m = F(p)
    loss = ((s - spec.to(DEV)) ** 2).mean() + 0.5 * ((m - mnorm.to(DEV)) ** 2).mean()
    loss.backward()
    assert all(q.grad is None for q in F.parameters())
    pr = pnorm.clone().requires_grad_(True)
    sr, mr = O.forward_model_forward(f_sd, pr)
    (((sr - spec) ** 2).mean() + 0.5 * ((mr - mnorm) ** 2).mean()).backward()
    assert rel(p.grad, pr.grad) < 5e-2, rel(p.grad, pr.grad)
