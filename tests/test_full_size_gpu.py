"""Parity at BASELINE config-2 FULL size (64 frames per GPU, 3 pairs at 320x240 / 160x120 / 80x60), where the CPU oracle
would take hours: size-independent properties of the operators, checked through the C ABI in the default BF16X3 mode.

* forward conv is affine in the frames: conv(a x1 + b x2) - bias = a (conv(x1) - bias) + b (conv(x2) - bias);
* the raw gradient block [GC | GF | GB | GP | sum e^2] is a SUM over frames: block(64 frames) = block(first 32) +
  block(last 32) -- this exercises the item / band / strip scheduling of the streaming kernels at the sizes bench.py
  measures (several items per CTA, all CTAs busy), against the same kernels on smaller launches that the oracle tests pin;
* repeated launches are bit-identical (fixed reduction order).
Tolerances: 1e-4 relative L2 (the fp32 bar of north_star); bit-exact for the determinism check."""
import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O

pytestmark = pytest.mark.gpu

PAIRS = [(16, 3, 320, 240), (32, 16, 160, 120), (64, 32, 80, 60)]  # (dM, dD, Nx, Ny), 5x5 taps
B = 64


@pytest.fixture()
def tc(ctx):
    ctx.set_precision(A.PRECISION_BF16X3)
    yield ctx
    ctx.set_precision(A.PRECISION_FP32)


def weights(rng, dM, dD):
    c = ((rng.random((dM, dD, 5, 5)) * 2 - 1) * 0.2).astype(np.float32)
    f = ((rng.random((dD, dM, 5, 5)) * 2 - 1) * 0.2).astype(np.float32)
    return c, f


@pytest.mark.parametrize("pair", PAIRS)
def test_conv_is_affine_in_the_frames_at_full_size(tc, pair):
    dM, dD, Nx, Ny = pair
    rng = np.random.default_rng(70)
    x1 = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    x2 = (rng.standard_normal((B, dD, Nx, Ny)) * 60).astype(np.float32)
    c, _ = weights(rng, dM, dD)
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    zero = np.zeros(dM, np.float32)
    y1 = tc.conv_fwd(x1, c, zero)
    y2 = tc.conv_fwd(x2, c, zero)
    y = tc.conv_fwd((0.5 * x1 - 2.0 * x2).astype(np.float32), c, b)
    want = 0.5 * y1 - 2.0 * y2 + b[None, :, None, None]
    assert O.rel_l2(y, want) < 1e-4


def block(ctx, dims, nB, inp, out, hin, c, f):
    dM, dD, Nx, Ny = dims
    n = int(A.lib().aefft_coord_gbuf_len(A.MODE_CUDA_REF, dD, dM, 5, 5))
    dev = [ctx.to_device(a) for a in (inp, out, hin, c, f)]
    g = A.DevBuf(ctx, (n,))
    ctx.coord_gradients(A.MODE_CUDA_REF, 0, nB, dD, dM, Nx, Ny, 5, 5, dev[0], dev[1], dev[2], dev[3], dev[4], g)
    ctx.sync()
    res = g.numpy().astype(np.float64)
    for d in dev:
        d.free()
    g.free()
    return res


@pytest.mark.parametrize("pair", PAIRS)
def test_gradient_block_adds_over_frames_at_full_size(tc, pair):
    dM, dD, Nx, Ny = pair
    rng = np.random.default_rng(71)
    inp = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    out = (inp + rng.standard_normal((B, dD, Nx, Ny)) * 40).astype(np.float32)
    hin = (rng.standard_normal((B, dM, Nx, Ny)) * 90 + 30).astype(np.float32)
    c, f = weights(rng, dM, dD)
    full = block(tc, pair, B, inp, out, hin, c, f)
    again = block(tc, pair, B, inp, out, hin, c, f)
    assert np.array_equal(full, again), "the gradient block must be bit-identical between launches"
    h = B // 2
    halves = block(tc, pair, h, inp[:h], out[:h], hin[:h], c, f) + block(tc, pair, h, inp[h:], out[h:], hin[h:], c, f)
    nC = dM * dD * 25
    for name, sl in (("GC", slice(0, nC)), ("GF", slice(nC, 2 * nC)), ("SQ", slice(2 * nC + dM + dD, 2 * nC + dM + dD + 1))):
        assert np.linalg.norm(full[sl] - halves[sl]) / np.linalg.norm(halves[sl]) < 1e-4, name
    # the bias sums cancel heavily: judge them against the magnitude of their terms
    gb, gp = slice(2 * nC, 2 * nC + dM), slice(2 * nC + dM, 2 * nC + dM + dD)
    assert np.abs(full[gb] - halves[gb]).max() < 1e-5 * np.abs(halves[gb]).max() + 1e-6 * np.abs(hin).sum() / dM
    assert np.abs(full[gp] - halves[gp]).max() < 1e-6 * np.abs(out - inp).sum() / dD
