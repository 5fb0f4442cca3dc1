"""Development aid: time the batched 2-D R2C / C2R transforms at the BASELINE config-3 sizes (frames resident)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "autoencoder-fft_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import aefft_ctypes as A

ctx = A.Ctx(0)
rng = np.random.default_rng(0)
for batch, N in [(384, 512), (512, 256), (1024, 128), (48, 1024)]:
    x = rng.standard_normal((batch, N, N)).astype(np.float32)
    xd = ctx.to_device(x)
    sp = A.DevBuf(ctx, (batch, N, N // 2 + 1, 2))
    back = A.DevBuf(ctx, (batch, N, N))
    for it in range(3):
        ctx.profile_enable(it == 2)
        ctx.fft_r2c(xd, loc=A.DEVICE, out=sp)
        ctx.fft_c2r(sp, N, loc=A.DEVICE, out=back)
    rows = ctx.profile_records()
    ctx.profile_enable(False)
    err = np.abs(back.numpy()[:2] / (N * N) - x[:2]).max()
    print(f"{batch} x {N}^2: " + ", ".join(f"{r['name']} {r['ms']:.3f} ms ({r['bytes'] / r['ms'] / 1e6:.0f} GB/s)" for r in rows) + f"  roundtrip err {err:.2e}", flush=True)
    for d in (xd, sp, back):
        d.free()
