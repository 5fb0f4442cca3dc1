"""Development probe: one backprop_fft call with the multiobjective term at the widest c4 pair (128 -> 256 channels: 32 768
kernels per tensor) on a tiny spatial size, so that gradient_diff_tiled_kernel dominates.  Prints its per-launch time; under
ncu (-k regex:gradient_diff_tiled) it is the capture target."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "autoencoder-fft_b200"))
import aefft_ctypes as A  # noqa: E402

dM, dD = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 128)
rng = np.random.default_rng(1)
ctx = A.Ctx(0)
c = (rng.standard_normal((dM, dD, 5, 5)) * 0.1).astype(np.float32)
f = (rng.standard_normal((dD, dM, 5, 5)) * 0.1).astype(np.float32)
b = (rng.standard_normal(dM) * 0.1).astype(np.float32)
p = (rng.standard_normal(dD) * 0.1).astype(np.float32)
x = rng.standard_normal((1, dD, 8, 8)).astype(np.float32)
o = rng.standard_normal((1, dD, 8, 8)).astype(np.float32)
for it in range(2):
    ctx.profile_enable(True)
    ctx.backprop_fft(x, x, o, c, f, b, p, 0.2, 1, 3)
    recs = {r["name"]: r for r in ctx.profile_records()}
    ctx.profile_enable(False)
g = recs["gradient_diff"]
print("gradient_diff: %.3f ms per launch (%d launches), %.1f TFLOP/s" % (g["ms"] / g["launches"], g["launches"], g["flops"] / g["ms"] / 1e9))
