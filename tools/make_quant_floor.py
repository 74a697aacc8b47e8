"""Writes tests/golden/quantisation_floor.json (CPU, float64; a few minutes): per-tensor relative distance between the
gradients of the exact float64 PI-GAN step and of the same step with the forward values rounded to fp16 where the
engine rounds them (oracle/quantised.py), at the batch sizes the GPU parity tests use.

    python tools/make_quant_floor.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import fixtures, quantised as Q  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
g_sd, d_sd, f_sd = fixtures.make_weights(42)
out = {"_doc": "relative L2 distance fp16-forward vs exact float64 gradients; seeds: weights 42, batch 100 "
               "(tests/test_gpu_engine.py::test_train_step_gradients_match_oracle)"}
for n in (4096, 16384, 65536):
    spec, praw, pnorm, mnorm = fixtures.make_batch(n, seed=100)
    t = time.time()
    out[str(n)] = Q.relative_floor(g_sd, d_sd, f_sd, (spec, praw, pnorm, None, mnorm))
    print(n, f"{time.time() - t:.1f} s", json.dumps(out[str(n)]), flush=True)
path = os.path.join(ROOT, "tests", "golden", "quantisation_floor.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
