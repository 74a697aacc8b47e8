"""The drop-in modules under autograd (torch.autograd.Function shims): the reference's UNFUSED training
iteration - D-step then G-step written with plain module calls, BCELoss and .backward(), as train_pigan.py:123-187 and
every variant trainer do - run through Generator / Discriminator / ForwardModel of this package on the GPU, against
the same iteration through the oracle's fp32 modules on the CPU.  Bounds: the quantisation-aware floor of
tests/golden/quantisation_floor.json applies to the gradients below a BatchNorm exactly as in
test_gpu_engine.py::test_train_step_gradients_match_oracle (fp16 forward operands); everything else 2e-3."""
import copy
import os
import sys

import pytest
import torch
import torch.nn.functional as F

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


def _modules():
    from core.models.discriminator import Discriminator
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from oracle import fixtures
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    G, D, Fm = Generator(250, 4), Discriminator(250, 4), ForwardModel(4, 250, 8)
    G.load_state_dict(g_sd); D.load_state_dict(d_sd); Fm.load_state_dict(f_sd)
    return (G.to(DEV), D.to(DEV), Fm.to(DEV).eval()), (g_sd, d_sd, f_sd)


def _denorm(p):                       # data_loader.py:238-252 on tensors
    return (p + 1.0) / 2.0 * 0.6 + 2.2


@pytest.mark.parametrize("n", [512, 8192])
def test_discriminator_step_through_autograd(n):
    """loss_d = BCE(D(x, p_real), 0.9) + BCE(D(x, denorm(G(x)).detach()), 0.1); loss_d.backward()."""
    from oracle import fixtures
    from oracle import models as O
    (G, D, _), (g_sd, d_sd, _) = _modules()
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=55)
    x, pr = spec.to(DEV), praw.to(DEV)
    G.train(); D.train()
    with torch.no_grad():
        fake = _denorm(G(x))
    out_r, out_f = D(x, pr), D(x, fake)
    assert out_r.requires_grad and out_r.shape == (n, 1)
    loss = F.binary_cross_entropy(out_r, torch.full_like(out_r, 0.9)) + \
        F.binary_cross_entropy(out_f, torch.full_like(out_f, 0.1))
    loss.backward()
    # oracle
    d = O._leaf(copy.deepcopy(d_sd), O.D_TRAINABLE)
    with torch.no_grad():
        fake_o = O.denormalize_params(O.generator_forward(copy.deepcopy(g_sd), spec, training=True))
    lo = O.bce(O.discriminator_forward(d, spec, praw), torch.full((n, 1), 0.9)) + \
        O.bce(O.discriminator_forward(d, spec, fake_o), torch.full((n, 1), 0.1))
    lo.backward()
    assert abs(float(loss.detach()) - float(lo.detach())) <= 1e-3 * abs(float(lo.detach()))
    for name, p in D.named_parameters():
        assert p.grad is not None and rel(p.grad, d[name].grad) < 3e-3, (name, rel(p.grad, d[name].grad))


@pytest.mark.parametrize("n", [4096])
def test_generator_step_through_autograd(n):
    """loss_g = BCE(D(x, denorm(G(x))), 1) + 10 * MSE(F(G(x)).spectrum, x): the adversarial gradient reaches G through
    the discriminator's `params` input, the reconstruction gradient through the frozen surrogate's input."""
    import json
    from oracle import fixtures
    from oracle import models as O
    (G, D, Fm), (g_sd, d_sd, f_sd) = _modules()
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=55)
    x = spec.to(DEV)
    G.train(); D.train()
    p = G(x)
    assert p.requires_grad
    out = D(x, _denorm(p))
    rec, _ = Fm(p)
    loss = F.binary_cross_entropy(out, torch.ones_like(out)) + 10.0 * F.mse_loss(rec, x)
    loss.backward()
    # oracle
    g = O._leaf(copy.deepcopy(g_sd), O.G_TRAINABLE)
    po = O.generator_forward(g, spec, training=True)
    oo = O.discriminator_forward(copy.deepcopy(d_sd), spec, O.denormalize_params(po))
    ro, _ = O.forward_model_forward(f_sd, po)
    lo = O.bce(oo, torch.ones(n, 1)) + 10.0 * F.mse_loss(ro, spec)
    lo.backward()
    assert abs(float(loss.detach()) - float(lo.detach())) <= 2e-3 * abs(float(lo.detach()))
    floor = json.load(open(os.path.join(ROOT, "tests", "golden", "quantisation_floor.json")))[str(n)]["g"]
    for name, prm in G.named_parameters():
        ref = g[name].grad
        if name in ("main.0.bias", "main.3.bias"):       # in front of a train-mode BatchNorm: the true gradient is 0,
            w = dict(G.named_parameters())[name.replace("bias", "weight")].grad      # both sides hold rounding noise
            assert float(prm.grad.abs().max()) < 1e-3 * float(w.abs().max())
            continue
        # the reconstruction term runs through five LayerNorm layers in fp16 (A19: 2.7e-2 on its own); the mix of
        # both terms is bounded by the larger of that and the BatchNorm floor
        bound = max(3e-2, 1.5 * floor.get(name, 0.0))
        assert rel(prm.grad, ref) < bound, (name, rel(prm.grad, ref), bound)
    # D's parameters received gradients too (the reference does not freeze D in the G-step)
    assert all(q.grad is not None for q in D.parameters())


def test_running_statistics_and_eval_mode():
    """Train-mode forward under autograd updates the BatchNorm buffers once (like nn.BatchNorm1d), its backward does
    not; eval mode under autograd refuses loudly; no_grad calls are unchanged."""
    from oracle import fixtures
    (G, D, _), _ = _modules()
    spec, *_ = fixtures.make_batch(256, seed=3)
    x = spec.to(DEV)
    G.train()
    nbt0 = int(G.main[1].num_batches_tracked)
    p = G(x)
    rm = G.main[1].running_mean.clone()
    p.sum().backward()
    assert int(G.main[1].num_batches_tracked) == nbt0 + 1 and torch.equal(rm, G.main[1].running_mean)
    with torch.no_grad():
        q = G(x)
    assert not q.requires_grad and torch.allclose(q, p.detach(), atol=1e-6)
    G.eval()
    with pytest.raises(NotImplementedError):
        G(x)
    with torch.no_grad():
        assert G(x).shape == (256, 4)


def test_interleaved_forwards_and_ragged_sizes():
    """The backward passes recompute the forward from the saved inputs, so any call order works: two generator
    forwards on different batches (130 and 257 rows: ragged row tiles), a discriminator forward in between, then the
    backward passes in reverse order; every gradient equals the one of an isolated call."""
    from oracle import fixtures
    (G, D, _), _ = _modules()
    G.train(); D.train()
    xs = [fixtures.make_batch(n, seed=s)[0].to(DEV) for n, s in ((130, 1), (257, 2))]
    pr = fixtures.make_batch(130, seed=1)[1].to(DEV)
    ws = [torch.randn(x.shape[0], 4, device=DEV) for x in xs]
    wd = torch.randn(130, 1, device=DEV)

    def isolated_g(i):
        G.zero_grad()
        (G(xs[i]) * ws[i]).sum().backward()
        return [p.grad.clone() for p in G.parameters()]

    ref = [isolated_g(0), isolated_g(1)]
    D.zero_grad()
    (D(xs[0], pr) * wd).sum().backward()
    ref_d = [p.grad.clone() for p in D.parameters()]
    G.zero_grad(); D.zero_grad()
    p0 = G(xs[0])
    out = D(xs[0], pr)
    p1 = G(xs[1])
    (p1 * ws[1]).sum().backward()
    g1 = [p.grad.clone() for p in G.parameters()]
    G.zero_grad()
    (out * wd).sum().backward()
    (p0 * ws[0]).sum().backward()
    g0 = [p.grad.clone() for p in G.parameters()]
    for a, b in zip(g0, ref[0]):
        assert torch.equal(a, b)
    for a, b in zip(g1, ref[1]):
        assert torch.equal(a, b)
    for a, b in zip([p.grad for p in D.parameters()], ref_d):
        assert torch.equal(a, b)
    # gradients accumulate across backward calls like any autograd leaf
    G.zero_grad()
    (G(xs[0]) * ws[0]).sum().backward()
    (G(xs[0]) * ws[0]).sum().backward()
    for p, b in zip(G.parameters(), ref[0]):
        assert torch.allclose(p.grad, 2 * b, rtol=1e-6, atol=0)


def test_backward_entry_points_validate_their_arguments():
    from pigan_b200 import engine as E
    from pigan_b200 import native
    eng = E.Engine(256, torch.device(DEV))
    g = torch.zeros(native.lib.pigan_generator_param_count(None), device=DEV)
    d = torch.zeros(native.lib.pigan_discriminator_param_count(None), device=DEV)
    x = torch.zeros(128, 250, device=DEV)
    gp = torch.zeros(128, 4, device=DEV)
    st = native.current_stream()
    bad = [
        lambda: native.lib.pigan_generator_backward(eng.handle, g.data_ptr(), x.data_ptr(), 1, gp.data_ptr(), 1.0, g.data_ptr(), st),
        lambda: native.lib.pigan_generator_backward(eng.handle, g.data_ptr(), x.data_ptr(), 128, gp.data_ptr(), 0.0, g.data_ptr(), st),
        lambda: native.lib.pigan_generator_backward(eng.handle, g.data_ptr(), None, 128, gp.data_ptr(), 1.0, g.data_ptr(), st),
        lambda: native.lib.pigan_generator_backward(eng.handle, g.data_ptr(), x.data_ptr(), 257, gp.data_ptr(), 1.0, g.data_ptr(), st),
        lambda: native.lib.pigan_discriminator_backward(eng.handle, d.data_ptr(), x.data_ptr(), gp.data_ptr(), 0, gp.data_ptr(), 1.0, d.data_ptr(), None, st),
        lambda: native.lib.pigan_discriminator_backward(eng.handle, d.data_ptr(), x.data_ptr(), None, 128, gp.data_ptr(), 1.0, d.data_ptr(), None, st),
        lambda: native.lib.pigan_discriminator_backward(eng.handle, d.data_ptr(), x.data_ptr(), gp.data_ptr(), 128, gp.data_ptr(), -1.0, d.data_ptr(), None, st),
    ]
    for f in bad:
        assert f() != native.PIGAN_OK and native.last_error()
    torch.cuda.synchronize()
