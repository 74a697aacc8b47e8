"""K7 physics kernel vs the C/NumPy oracle (oracle/physics.*): peak indices bit-exact, metrics 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FLOAT_RTOL = 1e-5  # north_star tolerance for fp32-accurate paths


def _run(spec, freq, peak_idx=None, baseline=0.0):
    from pigan_b200 import native
    n, s = spec.shape
    out_idx = torch.full((n,), -7, device="cuda", dtype=torch.int32)
    out = torch.zeros(n, 4, device="cuda", dtype=torch.float32)
    native.check(native.lib.pigan_physics_metrics(spec.data_ptr(), n, s, freq.data_ptr(),
                                                  None if peak_idx is None else peak_idx.data_ptr(),
                                                  float(baseline), out_idx.data_ptr(), out.data_ptr(),
                                                  native.current_stream()))
    torch.cuda.synchronize()
    return out_idx.cpu().numpy(), out.cpu().numpy()


def _compare(idx, out, ridx, rout):
    assert np.array_equal(idx, ridx)
    nan_g, nan_r = np.isnan(out), np.isnan(rout)
    assert np.array_equal(nan_g, nan_r)
    ok = ~nan_r
    np.testing.assert_allclose(out[ok], rout[ok].astype(np.float32), rtol=FLOAT_RTOL, atol=0)


@pytest.mark.parametrize("n,s", [(1, 250), (7, 250), (4096, 250), (100000, 250), (513, 501), (300, 2048), (33, 31)])
def test_physics_matches_oracle(n, s):
    from oracle import physics as P
    from pigan_b200 import synthetic
    spec, _, _, _ = synthetic.make_batch(n, s, seed=7 + n, device="cuda")
    freq = synthetic.frequencies(s, device="cuda")
    idx, out = _run(spec, freq)
    ridx, rout = P.physics_batch(spec.cpu().numpy(), freq.cpu().numpy())
    _compare(idx, out, ridx, rout)


def test_physics_given_peak_index_and_baseline():
    from oracle import physics as P
    from pigan_b200 import synthetic
    n, s = 5000, 250
    spec, _, _, _ = synthetic.make_batch(n, s, seed=3, device="cuda")
    freq = synthetic.frequencies(s, device="cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    pk = torch.randint(0, s, (n,), generator=g, dtype=torch.int32)
    idx, out = _run(spec, freq, pk.cuda(), baseline=-0.5)
    ridx, rout = P.physics_batch(spec.cpu().numpy(), freq.cpu().numpy(), pk.numpy(), baseline=-0.5)
    _compare(idx, out, ridx, rout)


def test_physics_edge_cases():
    """Ties (first occurrence), flat spectra, peak at the borders, equal neighbours, NaN outputs."""
    from oracle import physics as P
    from pigan_b200 import synthetic
    s = 250
    freq = synthetic.frequencies(s, device="cuda")
    rows = []
    rows.append(np.zeros(s, np.float32))                                   # flat: no crossing -> NaN
    r = np.zeros(s, np.float32); r[0] = -5; rows.append(r)                 # peak at first sample
    r = np.zeros(s, np.float32); r[-1] = -5; rows.append(r)                # peak at last sample
    r = np.zeros(s, np.float32); r[100] = -4; r[150] = -4; rows.append(r)  # tie -> first
    r = -np.abs(np.linspace(-1, 1, s)).astype(np.float32); rows.append(r)  # V shape, minimum at the ends
    r = -np.ones(s, np.float32); r[120:130] = -3; rows.append(r)           # plateau minimum + equal neighbours
    r = np.zeros(s, np.float32); r[60] = -10; r[59] = -5; r[61] = -5; rows.append(r)  # crossing exactly on samples
    r = np.full(s, -1e-7, np.float32); r[10] = -5e-7; rows.append(r)       # |t_min| < 1e-6 -> FoM NaN
    spec = torch.from_numpy(np.stack(rows)).cuda()
    idx, out = _run(spec, freq)
    ridx, rout = P.physics_batch(spec.cpu().numpy(), freq.cpu().numpy())
    _compare(idx, out, ridx, rout)
    # the scalar Python restatement agrees with the C one on the same rows
    pidx, pout = P.physics_rows_python(spec.cpu().numpy(), freq.cpu().numpy())
    assert np.array_equal(pidx, ridx)
    np.testing.assert_array_equal(np.isnan(pout), np.isnan(rout))
    np.testing.assert_allclose(pout[~np.isnan(pout)], rout[~np.isnan(rout)], rtol=1e-14)


def test_physics_empty():
    from pigan_b200 import native
    assert native.lib.pigan_physics_metrics(None, 0, 250, None, None, 0.0, None, None, None) == 0


def test_physics_sharded_by_spectrum_equals_unsharded():
    """the kernel shards by spectrum with no collective - the ranks' results (physics.shard_rows ranges)
    concatenate to the single-GPU result bit for bit, for any world size."""
    from oracle import fixtures
    from pigan_b200 import physics
    n = 10007
    spec, _, _, _ = fixtures.make_batch(n, seed=91)
    x = spec.cuda()
    full = physics.peak_metrics(x)
    for world in (2, 3, 8):
        parts = [physics.peak_metrics(x[slice(*physics.shard_rows(n, r, world))]) for r in range(world)]
        for k in ("peak_idx", "f_res", "Q", "FoM", "S"):
            cat = torch.cat([p[k] for p in parts])
            assert torch.equal(torch.nan_to_num(cat.float(), nan=-7.0), torch.nan_to_num(full[k].float(), nan=-7.0)), (world, k)
    s = physics.sharded_peak_summary(full)
    assert s["spectra"] == n and 0 < s["defined_Q"] <= n
