"""CPU proof behind the gradient tolerances of the GPU parity tests (BASELINE north_star: gradients within 1e-3).

oracle/quantised.py evaluates the PI-GAN step in float64 with the forward values rounded to fp16 exactly where the
engine rounds them.  The distance of its gradients from the exact float64 step is what ANY fp16-operand
implementation shows against the reference, however exact its arithmetic.  These tests pin that floor:
the discriminator and the generator layers above the last BatchNorm stay below 1e-3; the generator's lower layers do
not - and every rounding site ALONE (input operand, weight copies, stored activations) already exceeds 1e-3 there, so
neither a hi+lo split of the spectrum operand nor fp32 activations would bring them under the stated tolerance.
tests/test_gpu_engine.py::test_train_step_gradients_match_oracle checks the engine (a) against the plain oracle at
1.5 x this floor and (b) against the fp16-forward restatement itself at 1e-3."""
import copy
import json
import os

import torch

from oracle import fixtures
from oracle import models as O
from oracle import quantised as Q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOOR = json.load(open(os.path.join(ROOT, "tests", "golden", "quantisation_floor.json")))
LOWER = ("main.0.weight", "main.1.weight", "main.1.bias", "main.3.weight")      # below the last BatchNorm
UPPER = ("main.4.weight", "main.4.bias", "main.6.weight", "main.6.bias")


def _batch(n):
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=100)
    return spec, praw, pnorm, None, mnorm


def test_exact_restatement_equals_the_oracle_step():
    """With the rounding switched off the restatement is train_pigan.py:123-187 itself: it reproduces
    oracle.models.train_step (pinned to the reference by tests/golden) to float64 round-off."""
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    batch = _batch(512)
    de, ge = Q.train_step_grads(g_sd, d_sd, f_sd, batch, quantise=False)
    g2, d2, f2 = (O.cast_state(copy.deepcopy(s), torch.float64) for s in (g_sd, d_sd, f_sd))
    b64 = tuple(None if x is None else x.double() for x in batch)
    _, ex = O.train_step(g2, d2, f2, O.Adam(O.G_TRAINABLE), O.Adam(O.D_TRAINABLE), b64, 2e-4, 2e-4)
    for k, v in de.items():
        assert float((v - ex["d_grads"][k]).norm() / ex["d_grads"][k].norm()) < 1e-10, k
    for k, v in ge.items():
        if k in ("main.0.bias", "main.3.bias"):      # exactly zero here, float64 round-off in autograd's BatchNorm
            assert float(ex["g_grads"][k].norm()) < 1e-12 * float(ex["g_grads"]["main.0.weight"].norm()) + 1e-15
            continue
        assert float((v - ex["g_grads"][k]).norm() / ex["g_grads"][k].norm()) < 1e-10, k


def test_committed_floor_is_current_and_has_the_documented_shape():
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    Q.set_sites(Q.SITES)
    fl = Q.relative_floor(g_sd, d_sd, f_sd, _batch(4096))
    for net in ("d", "g"):
        for k, v in fl[net].items():
            ref = FLOOR["4096"][net][k]
            assert abs(v - ref) <= 0.02 * ref + 1e-12, (net, k, v, ref)   # tools/make_quant_floor.py must be re-run
    # shape of the floor at every batch size of the GPU test
    for n in ("4096", "16384", "65536"):
        assert max(FLOOR[n]["d"].values()) < 1.2e-3
        assert all(FLOOR[n]["g"][k] < 1e-3 for k in UPPER)
        assert all(FLOOR[n]["g"][k] > 1.2e-3 for k in LOWER)
    # it shrinks with the batch size (the rounding errors of the rows average out)
    assert FLOOR["65536"]["g"]["main.0.weight"] < 0.5 * FLOOR["4096"]["g"]["main.0.weight"]


def test_each_rounding_site_alone_exceeds_the_stated_tolerance_below_the_last_batchnorm():
    """Ablation at batch 4096: rounding only the input operand, only the weight copies or only the stored
    activations each moves main.0.weight's gradient by more than 1e-3 - no single mitigation closes the gap."""
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd, d_sd, f_sd = fixtures.make_weights(42)
    batch = _batch(4096)
    try:
        for site in Q.SITES:
            Q.set_sites((site,))
            fl = Q.relative_floor(g_sd, d_sd, f_sd, batch)
            assert fl["g"]["main.0.weight"] > 2e-3, (site, fl["g"]["main.0.weight"])
            assert max(fl["g"][k] for k in UPPER) < 1e-3, site
    finally:
        Q.set_sites(Q.SITES)


def _surrogate_input_grad(f64, pnorm, spec, on, stored_pre_ln=False):
    p = pnorm.double().clone().requires_grad_(True)
    s, _ = Q.forward_model_forward(f64, p, spec.shape[1], on, stored_pre_ln=stored_pre_ln)
    ((s - spec.double()) ** 2).mean().backward()
    return p.grad


def test_surrogate_input_gradient_floor():
    """SURVEY A19 (d MSE(F(p).spectrum, x)/dp with F frozen): rounding the forward values of the surrogate to fp16
    where the engine does - weight copies from layer 2 on, activations between the layers, and (backward-capable
    path) the stored Linear outputs in front of each LayerNorm - moves the float64 input gradient by 2.7-3.0e-2; the
    weights alone or the activations alone by more than 1e-2.  The gradient passes five LayerNorm backward
    projections and ends in a 4-wide layer: it is a small difference of large terms.  The GPU test
    (tests/test_gpu_fwd_train.py) measures 2.7e-2 against the fp32 oracle, i.e. the floor itself, and checks the
    engine against THIS restatement: 4e-3 ... 1.6e-2, the part its fp16 backward tensors add."""
    torch.set_num_threads(os.cpu_count() or 1)
    _, _, f_sd = fixtures.make_weights(42)
    f64 = O.cast_state(copy.deepcopy(f_sd), torch.float64)
    try:
        for n, lo, hi in ((256, 2.0e-2, 3.2e-2), (4096, 2.2e-2, 3.2e-2)):
            spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=23)
            Q.set_sites(Q.SITES)
            exact = _surrogate_input_grad(f64, pnorm, spec, False)
            fl = float((_surrogate_input_grad(f64, pnorm, spec, True) - exact).norm() / exact.norm())
            assert lo < fl < hi, (n, fl)
            fl2 = float((_surrogate_input_grad(f64, pnorm, spec, True, True) - exact).norm() / exact.norm())
            assert lo < fl2 < hi, (n, fl2)
            for site in ("weights", "activations"):
                Q.set_sites((site,))
                one = float((_surrogate_input_grad(f64, pnorm, spec, True) - exact).norm() / exact.norm())
                assert one > 1e-2, (n, site, one)
    finally:
        Q.set_sites(Q.SITES)
