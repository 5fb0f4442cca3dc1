"""Development aid: time the coordinate-space gradient kernels of each BASELINE config-2 pair in isolation
(B frames resident), optionally with the role wait counters of wgrad_ts (AEFFT_TS_DEBUG=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "autoencoder-fft_b200"))
import aefft_ctypes as A

B = int(os.environ.get("B", 64))
ctx = A.Ctx(0)
ctx.set_precision(A.PRECISION_BF16 if os.environ.get("PREC") == "bf16" else A.PRECISION_BF16X3)
rng = np.random.default_rng(0)
for dM, dD, Nx, Ny in [(16, 3, 320, 240), (32, 16, 160, 120), (64, 32, 80, 60)]:
    inp = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    out = (inp + rng.standard_normal((B, dD, Nx, Ny)).astype(np.float32) * 40)
    hin = (rng.standard_normal((B, dM, Nx, Ny)) * 90 + 30).astype(np.float32)
    c = ((rng.random((dM, dD, 5, 5)) * 2 - 1) * 0.2).astype(np.float32)
    f = ((rng.random((dD, dM, 5, 5)) * 2 - 1) * 0.2).astype(np.float32)
    n = int(A.lib().aefft_coord_gbuf_len(A.MODE_CUDA_REF_SYM, dD, dM, 5, 5))
    dev = [ctx.to_device(a) for a in (inp, out, hin, c, f)]
    g = A.DevBuf(ctx, (n,))
    for it in range(3):
        if it == 2:
            ctx.profile_enable(True)
        ctx.coord_gradients(A.MODE_CUDA_REF_SYM, 0, B, dD, dM, Nx, Ny, 5, 5, *dev, g)
    ctx.sync()
    rows = ctx.profile_records()
    ctx.profile_enable(False)
    print(f"pair {dD}->{dM} {Nx}x{Ny} B={B}: " + ", ".join(f"{r['name']} {r['ms']:.3f} ms" for r in rows), flush=True)
    for d in dev:
        d.free()
    g.free()

# forward / decoder convolutions of each pair (host arrays in, only the kernels are timed)
for dM, dD, Nx, Ny in [(16, 3, 320, 240), (32, 16, 160, 120), (64, 32, 80, 60)]:
    x = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, 5, 5)) * 2 - 1) * 0.2).astype(np.float32)
    f = ((rng.random((dD, dM, 5, 5)) * 2 - 1) * 0.2).astype(np.float32)
    b = np.zeros(dM, np.float32)
    pb = np.zeros(dD, np.float32)
    for it in range(2):
        ctx.profile_enable(it == 1)
        h = ctx.conv_fwd(x, c, b)
        y = ctx.conv_fwd(h, f, pb)
    rows = ctx.profile_records()
    ctx.profile_enable(False)
    print(f"conv pair {dD}->{dM}->{dD} {Nx}x{Ny} B={B}: " + ", ".join(f"{r['name']} {r['ms']:.3f} ms/{r['launches']}" for r in rows), flush=True)
