"""Generates tests/golden/ctf_golden.npz by running the UNMODIFIED reference functions (src/ctf.py `ctf_filter`,
src/image.py `crop`, the --normalize arithmetic of train_particles.py:592-600) in the build container.

    python oracle/make_golden_ctf.py        # needs /root/reference; the fixture is committed, the reference is not
"""
import os
import sys

import numpy as np
import pandas as pd

REF = os.environ.get("TVAE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import src.ctf as C          # noqa: E402
import src.image as I        # noqa: E402

rng = np.random.default_rng(77)
B = 5
params = np.stack([rng.uniform(1.0, 3.0, B), np.full(B, 2.7), np.full(B, 300.0), rng.choice([1.3, 2.6], B),
                   rng.choice([0.0, 100.0, 250.0], B), np.full(B, 10.0), np.zeros(B), rng.uniform(0, 180, B)], 1)
df = pd.DataFrame(params, columns=['defocus', 'cs', 'voltage', 'apix', 'bfactor', 'ampcont', 'dfdiff', 'dfang'])
out = {"params": params}
for n, m, scale in ((15, 15, 1), (31, 33, 1), (32, 20, 2), (127, 127, 1)):
    take = df.iloc[:2].reset_index(drop=True) if n == 127 else df
    out[f"ctf_{n}_{m}_{scale}"] = C.ctf_filter(take, n, m, scale=scale)
stack = (rng.standard_normal((4, 40, 36)) * 3.0 + 1.5).astype(np.float32)
out["stack"] = stack
c = I.crop(stack, 24)
mu = c.reshape(-1, 24 * 24).mean(1)
std = c.reshape(-1, 24 * 24).std(1)
out["crop24_norm"] = (c - mu[:, np.newaxis, np.newaxis]) / std[:, np.newaxis, np.newaxis]
mu = stack.reshape(-1, 40 * 36).mean(1)
std = stack.reshape(-1, 40 * 36).std(1)
out["norm"] = (stack - mu[:, np.newaxis, np.newaxis]) / std[:, np.newaxis, np.newaxis]
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ctf_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
