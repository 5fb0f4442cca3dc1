"""GPU parity tests of the streaming TMEM-operand weight-gradient kernel (wgrad_ts.cu): the raw gradient block
[GC | GF | GB | GP | sum e^2] of aefft_coord_gradients in BF16X3 precision against (a) the numpy oracle on small
frames and (b) the engine's own fp32 CUDA-core kernels (themselves pinned to the oracle in test_coord_gpu.py) at the
BASELINE config-2 shapes, where the oracle would take minutes.  Tolerance: 1e-4 relative L2 per block (fp32 bar)."""
import os

import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O
from test_coord_gpu import make_case

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc(ctx):
    yield ctx
    ctx.set_precision(A.PRECISION_FP32)


def gradient_block(ctx, precision, dims, B, inp, out, hin, c, f):
    dM, dD, Nk, Nl, Nx, Ny = dims
    n = int(A.lib().aefft_coord_gbuf_len(A.MODE_CUDA_REF, dD, dM, Nk, Nl))
    ctx.set_precision(precision)
    dev = [ctx.to_device(a) for a in (inp, out, hin, c, f)]
    g = A.DevBuf(ctx, (n,))
    ctx.profile_enable(True)
    ctx.coord_gradients(A.MODE_CUDA_REF, 0, B, dD, dM, Nx, Ny, Nk, Nl, dev[0], dev[1], dev[2], dev[3], dev[4], g)
    ctx.sync()
    names = [r["name"] for r in ctx.profile_records()]
    ctx.profile_enable(False)
    res = g.numpy()
    for d in dev:
        d.free()
    g.free()
    nC = dM * dD * Nk * Nl
    blocks = dict(GC=res[:nC], GF=res[nC:2 * nC], GB=res[2 * nC:2 * nC + dM], GP=res[2 * nC + dM:2 * nC + dM + dD],
                  SQ=res[2 * nC + dM + dD:2 * nC + dM + dD + 1])
    return blocks, names


def random_case(seed, dims, B):
    """Inputs shaped like a real step: integer pixel frames, hidden maps and reconstructions of matching magnitude
    (the gradients are a function of (in, out, hin, f) only; they need not come from a forward pass)."""
    dM, dD, Nk, Nl, Nx, Ny = dims
    rng = np.random.default_rng(seed)
    inp = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    out = (inp + rng.standard_normal((B, dD, Nx, Ny)) * 40).astype(np.float32)
    hin = (rng.standard_normal((B, dM, Nx, Ny)) * 90 + 30).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    f = ((rng.random((dD, dM, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    return inp, out, hin, c, f


SMALL = [(16, 3, 5, 5, 40, 48), (32, 16, 5, 5, 33, 60), (64, 32, 5, 5, 24, 20), (8, 5, 3, 3, 30, 16), (128, 4, 5, 5, 20, 24),
         (16, 3, 5, 5, 37, 132)]


@pytest.mark.parametrize("asmem", ["auto", "0", "2"])
@pytest.mark.parametrize("stack", ["auto", "0", "1"])
@pytest.mark.parametrize("dims", SMALL)
def test_gradient_block_vs_oracle(tc, dims, stack, asmem, monkeypatch):
    """`stack` selects the B operand form of wgrad_ts (separate hi / lo planes or the stacked swizzled plane), `asmem` the
    home of the A operand (tensor memory, or the shared-memory row ring; "2" also allows its 64-pixel strips); the
    defaults pick per layer shape, so every form is forced here for every shape."""
    dM, dD, Nk, Nl, Nx, Ny = dims
    B = 3
    if asmem == "auto":
        monkeypatch.delenv("AEFFT_TS_ASMEM", raising=False)
    else:
        monkeypatch.setenv("AEFFT_TS_ASMEM", asmem)
    if stack == "auto":
        monkeypatch.delenv("AEFFT_TS_STACK", raising=False)
    else:
        monkeypatch.setenv("AEFFT_TS_STACK", stack)
    cs = make_case(31, *dims, B=B)
    got, names = gradient_block(tc, A.PRECISION_BF16X3, dims, B, cs["inp"], cs["out"], cs["hin"], cs["c"], cs["f"])
    assert "wgrad_ts" in names, f"streaming kernel not used: {names}"
    norm = float(np.float32(dD * dM * Nk * Nl * Nx * Ny))
    want = dict(GC=0, GF=0, GB=0, GP=0, SQ=0)
    for n in range(B):
        gC, gF, gB, gP, mse = O.coord_gradients_cuda(cs["inp"][n], cs["out"][n], cs["hin"][n], cs["c"], cs["f"], False,
                                                     quirks=False)
        want["GC"] = want["GC"] + gC * norm
        want["GF"] = want["GF"] + gF * norm
        want["GB"] = want["GB"] + gB * norm
        want["GP"] = want["GP"] + gP * norm
        want["SQ"] = want["SQ"] + mse * norm
    for k in ("GC", "GF", "GB", "GP", "SQ"):
        w = np.asarray(want[k], np.float64).reshape(-1)
        scale = np.linalg.norm(w) + 1e-30
        # the bias sums cancel almost completely (|sum| << sum|.|): judge them against the magnitude of their terms
        if k == "GB":
            scale = max(scale, 1e-3 * np.abs(cs["hin"]).sum() * 0.2)
        if k == "GP":
            scale = max(scale, 1e-6 * np.abs(cs["out"] - cs["inp"]).sum())
        assert np.linalg.norm(got[k].astype(np.float64) - w) / scale < 1e-4, k


C2_SHAPES = [((16, 3, 5, 5, 320, 240), 4), ((32, 16, 5, 5, 160, 120), 6), ((64, 32, 5, 5, 80, 60), 16)]


@pytest.mark.parametrize("asmem", ["auto", "0", "2"])
@pytest.mark.parametrize("dims,B", C2_SHAPES + [((32, 16, 5, 5, 160, 60), 2)])
def test_gradient_block_vs_fp32_kernels_at_config2_shapes(tc, dims, B, asmem, monkeypatch):
    if asmem == "auto":
        monkeypatch.delenv("AEFFT_TS_ASMEM", raising=False)
    else:
        monkeypatch.setenv("AEFFT_TS_ASMEM", asmem)
    inp, out, hin, c, f = random_case(32, dims, B)
    ref, names32 = gradient_block(tc, A.PRECISION_FP32, dims, B, inp, out, hin, c, f)
    assert "wgrad_ts" not in names32
    got, names = gradient_block(tc, A.PRECISION_BF16X3, dims, B, inp, out, hin, c, f)
    assert "wgrad_ts" in names, f"streaming kernel not used: {names}"
    for k in ("GC", "GF", "SQ"):
        assert O.rel_l2(got[k], ref[k]) < 1e-4, k
    # sums with cancellation: compare against the magnitude of the summed terms
    assert np.abs(got["GP"] - ref["GP"]).max() < 1e-6 * np.abs(out - inp).sum() / dims[1]
    assert np.abs(got["GB"] - ref["GB"]).max() < 1e-5 * np.abs(ref["GB"]).max() + 1e-6 * np.abs(hin).sum() / dims[0]


def test_single_pass_bf16_is_the_looser_mode(tc):
    dims, B = (32, 16, 5, 5, 48, 60), 2
    inp, out, hin, c, f = random_case(33, dims, B)
    ref, _ = gradient_block(tc, A.PRECISION_FP32, dims, B, inp, out, hin, c, f)
    got, names = gradient_block(tc, A.PRECISION_BF16, dims, B, inp, out, hin, c, f)
    assert "wgrad_ts" in names
    r = O.rel_l2(got["GC"], ref["GC"])
    assert 1e-6 < r < 1e-2


def test_unsupported_shapes_fall_back(tc):
    dims, B = (8, 5, 3, 3, 33, 17), 2  # Ny not a multiple of 4: outside the TMA envelope
    inp, out, hin, c, f = random_case(34, dims, B)
    ref, _ = gradient_block(tc, A.PRECISION_FP32, dims, B, inp, out, hin, c, f)
    got, names = gradient_block(tc, A.PRECISION_BF16X3, dims, B, inp, out, hin, c, f)
    assert "wgrad_ts" not in names
    assert O.rel_l2(got["GC"], ref["GC"]) < 1e-4
