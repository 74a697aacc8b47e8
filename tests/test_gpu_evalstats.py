"""Device-side evaluator reductions (pigan_regression_* / pigan_score_summary_*) against the golden
values produced by the reference's own UnifiedEvaluator.calculate_metrics (sklearn + scipy) and numpy, and against
the oracle restatement (oracle/evalstats.py).  Sums are fp64: 1e-6 relative; MAPE replicates numpy's float32
division, 1e-5."""
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
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def _close(a, b, rtol):
    if np.isnan(b):
        assert np.isnan(a)
    else:
        assert abs(a - b) <= rtol * abs(b) + 1e-9, (a, b)


def test_regression_metrics_match_reference_golden_and_oracle():
    from oracle import evalstats as ES
    from oracle import fixtures
    from pigan_b200 import evalstats
    g = np.load(os.path.join(GOLD, "evaluator_metrics.npz"))
    cases, _ = fixtures.evaluator_cases()
    for name, (y, p) in cases.items():
        got = evalstats.regression_metrics(torch.from_numpy(y).to(DEV), torch.from_numpy(p).to(DEV))
        ref = ES.regression_metrics(y, p)
        for k in evalstats.REGRESSION_KEYS:
            _close(got[k], float(g[f"{name}_{k}"]), 2e-5)
            _close(got[k], ref[k], 1e-5 if k == "mape" else 1e-6)


def test_streaming_batches_equal_one_shot_and_ragged_sizes():
    from oracle import evalstats as ES
    from pigan_b200 import evalstats
    rng = np.random.Generator(np.random.PCG64(9))
    for n, c in ((1, 4), (7, 1), (1000, 250), (4097, 8), (129, 33)):
        y = rng.standard_normal((n, c)).astype(np.float32) * 3 - 1
        p = (y + 0.2 * rng.standard_normal((n, c))).astype(np.float32)
        yg, pg = torch.from_numpy(y).to(DEV), torch.from_numpy(p).to(DEV)
        one = evalstats.regression_metrics(yg, pg)
        m = evalstats.RegressionMetrics(c, DEV)
        for lo in range(0, n, 300):
            m.update(yg[lo:lo + 300], pg[lo:lo + 300])
        many, ref = m.compute(), ES.regression_metrics(y, p)
        for k in evalstats.REGRESSION_KEYS:
            if n == 1 and k in ("r2", "pearson_r"):
                continue    # a single row has no variance: sklearn warns / scipy raises, not part of the contract
            _close(one[k], ref[k], 1e-5)
            _close(many[k], ref[k], 1e-5)


def test_constant_columns_follow_sklearn_and_scipy():
    from pigan_b200 import evalstats
    y = torch.ones(64, 3, device=DEV)
    same, off = evalstats.regression_metrics(y, y), evalstats.regression_metrics(y, y + 1)
    assert same["r2"] == 1.0 and off["r2"] == 0.0 and np.isnan(same["pearson_r"]) and same["mse"] == 0.0
    assert abs(off["mape"] - 100.0) < 1e-4


def test_full_size_properties():
    """1M x 250 (1 GB per array): identities that hold at any size."""
    from pigan_b200 import evalstats
    g = torch.Generator(device=DEV).manual_seed(3)
    y = torch.randn(1 << 20, 250, device=DEV, generator=g)
    m = evalstats.regression_metrics(y, y)
    assert m["mse"] == 0.0 and m["r2"] == 1.0 and abs(m["pearson_r"] - 1.0) < 1e-9
    lin = evalstats.regression_metrics(y, 2.0 * y + 3.0)
    assert abs(lin["pearson_r"] - 1.0) < 1e-6
    p = y + 0.5
    m = evalstats.regression_metrics(y, p)
    assert abs(m["mse"] - 0.25) < 1e-6 and abs(m["mae"] - 0.5) < 1e-6 and abs(m["rmse"] - 0.5) < 1e-6
    ref_r2 = 1.0 - 0.25 / float(y.double().var(dim=0, unbiased=False).mean()) 
    assert abs(m["r2"] - (1.0 - float((0.25 / y.double().var(dim=0, unbiased=False)).mean()))) < 1e-6, ref_r2


def test_score_summary_matches_golden_and_scoring_outputs():
    from oracle import evalstats as ES
    from oracle import fixtures
    from pigan_b200 import evalstats
    g = np.load(os.path.join(GOLD, "evaluator_metrics.npz"))
    _, (viol, err, cons) = fixtures.evaluator_cases()
    s = evalstats.ScoreSummary(DEV)
    for lo in range(0, 1000, 256):
        s.update(torch.from_numpy(viol[lo:lo + 256]).to(DEV), torch.from_numpy(err[lo:lo + 256]).to(DEV),
                 torch.from_numpy(cons[lo:lo + 256]).to(DEV))
    got = s.compute()
    assert got["num_samples"] == 1000
    for k in evalstats.SUMMARY_KEYS:
        _close(got[k], float(g["summ_" + k]), 2e-6)
    # straight from the scoring kernel's outputs, nothing copied to the host in between
    from core.models.forward_model import ForwardModel
    from core.models.generator import Generator
    from pigan_b200 import engine as E
    from pigan_b200 import flat
    g_sd, _, f_sd = fixtures.make_weights(42)
    G, F = Generator(250, 4), ForwardModel(4, 250, 8)
    G.load_state_dict(g_sd); F.load_state_dict(f_sd)
    G.to(DEV).eval(); F.to(DEV).eval()
    eng = E.Engine(512, torch.device(DEV))
    eng.load_forward_model(flat.net_state(F, "forward_model").params.tensor())
    spec, _, _, _ = fixtures.make_batch(512, seed=21)
    st = flat.net_state(G, "generator")
    res = eng.score_candidates(st.params.tensor(), st.bn.tensor(), spectra=spec.to(DEV))
    s = evalstats.ScoreSummary(DEV)
    s.update(res["violations"], res["recon_error"], res["consistency"])
    got = s.compute()
    ref = ES.score_summary(res["violations"].cpu().numpy(), res["recon_error"].cpu().numpy(),
                           res["consistency"].cpu().numpy())
    for k in evalstats.SUMMARY_KEYS:
        _close(got[k], ref[k], 1e-6)
