"""TEST INFRASTRUCTURE — CPU restatement (numpy float64) of the spectrum part of
generate_single_terahertz_spectrum_and_params (core/utils/data_loader.py:62-80) with the noise passed in explicitly.
Pinned to the reference function by tests/golden/datagen.npz (tools/make_golden.py: datagen).  Only tests/ may
import this."""
import numpy as np


def generate_spectra(frequency, params, noise=None, noise_level=0.1, apply_offset=True):
    f = np.asarray(frequency, dtype=np.float64)[None, :]
    p = np.asarray(params, dtype=np.float64)
    r1, r2, w, g = (p[:, i:i + 1] - 2.5 for i in range(4))
    c1 = 0.870 + r1 * 0.05 + w * 0.03                      # :64
    m1 = -12.657 + r2 * 1.5 - g * 1.0                      # :65
    w1 = 0.08 + np.abs(r1 * 0.02)                          # :66
    t = m1 * np.exp(-((f - c1) ** 2) / (2 * w1 ** 2))      # :67
    c2 = 2.115 + r2 * 0.07 + g * 0.04                      # :69
    m2 = -11.763 + r1 * 1.0 - w * 0.8                      # :70
    w2 = 0.15 + np.abs(r2 * 0.03)                          # :71
    t = t + m2 * np.exp(-((f - c2) ** 2) / (2 * w2 ** 2))  # :72-73
    t = t + -0.5 * (np.tanh((f - 1.5) * 2) + 1)            # :74
    if apply_offset:
        t = t + (-0.5 + 0.5 * (f / 3.0))                   # :75-77
    if noise is not None:
        t = t + noise_level * np.asarray(noise, dtype=np.float64)   # :78-79
    return np.minimum(t, 0)                                # :80
