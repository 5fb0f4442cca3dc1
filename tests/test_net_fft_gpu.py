"""GPU parity tests of the momentum-space mode of the device-resident net (csrc/net_fft.cu): the forward that keeps every
layer's spectrum, the training step that consumes those spectra, and net_cfreq as a view of the resident kernels --
against the fp64 oracle (autoenc_fft + backprop_fft per pair, fft_backproplib.cu:1331-1511) and against this engine's
reference-shaped C-ABI path (aefft_autoenc_fft with fft_l = 1, then aefft_backprop_fft on the real-space layers), which
the other test files pin against the compiled reference.  Tolerances: layers 2e-5, weights 1e-4 / updates 2e-3 vs the
oracle (north_star), 2e-5 between the two engine paths (they differ by one C2R/R2C round trip in fp32)."""
import ctypes

import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O

pytestmark = pytest.mark.gpu


def make_net(ctx, D, Nx, Ny, widths, pools, B, rmax=0.1, seed=4321):
    ctypes.CDLL("libc.so.6").srand(seed)
    net = A.Net(ctx, D, Nx, Ny, B)
    for m, s in zip(widths, pools):
        net.add_layer(m, 1, 1, s, rmax)
    N = 2 * net.num_pairs
    convs = [net.get_conv(n) for n in range(N)]
    net_c, net_b = [c for c, _ in convs], [b for _, b in convs]
    scale = [net.conv_dims(n)[4] for n in range(N)]
    shapes = [net.layer_info(l)[:3] for l in range(net.num_layers)]
    return net, net_c, net_b, scale, shapes


CASES = [
    (3, 32, 32, [4, 5], [2, 2], 2),          # no tensor-core pair: the reference's bins-fastest layout throughout
    (3, 64, 64, [8, 16], [2, 2], 16),        # pair 0 CUDA cores, pair 1 tensor cores (layout change at the pooling)
    (8, 32, 64, [16, 8], [2, 1], 17),        # both pairs tensor cores, pool 1 between them, rectangular
    (3, 64, 128, [8, 16, 32], [2, 2, 2], 16), # the c3 stack at reduced size
    (3, 160, 120, [4, 8], [2, 2], 3),         # the camera's aspect (factors 3 and 5): 80 x 60 and 40 x 30 levels
    (3, 80, 120, [8, 16], [2, 2], 16),        # the same with a tensor-core pair (40 x 60, 20 x 30)
]


@pytest.mark.parametrize("cfg", CASES)
def test_net_fft_forward_vs_oracle_and_capi(ctx, cfg):
    D, Nx, Ny, widths, pools, B = cfg
    net, net_c, net_b, scale, shapes = make_net(ctx, D, Nx, Ny, widths, pools, B)
    try:
        x = O.synth_frames(7, B, D, Nx, Ny)
        net.fft_forward(x, fft_l=1)
        ctx.sync()
        got = [net.layer(l) for l in range(net.num_layers)]
        capi, spectra = ctx.autoenc_fft(x, net_c, net_b, scale, shapes, None, 1)
        for n in range(B):
            want, _ = O.autoenc_fft(x[n], net_c, net_b, scale, None, 1)
            for l in range(len(shapes)):
                assert O.rel_l2(got[l][n], want[l]) < 2e-5, (n, l)
        for l in range(len(shapes)):
            assert O.rel_l2(got[l], capi[l]) < 2e-5, l
        # fft_l = 0 writes the last layer only
        net.fft_forward(x * 0.5, fft_l=0)
        ctx.sync()
        assert O.rel_l2(net.layer(len(shapes) - 1), 0.5 * got[-1] + 0.5 * O.autoenc_fft(np.zeros_like(x[0]), net_c, net_b, scale, None, 0)[0][-1]) < 1e-4
        assert np.array_equal(net.layer(1), got[1])
        # net_cfreq as a view of the resident kernels
        for n in range(len(net_c)):
            assert O.rel_l2(net.get_cfreq(n).ravel(), spectra[n]) < 1e-5, n
    finally:
        net.close()


@pytest.mark.parametrize("cfg", CASES)
@pytest.mark.parametrize("maxdiff", [0, 1])
def test_net_fft_step_vs_oracle_and_capi(ctx, cfg, maxdiff):
    D, Nx, Ny, widths, pools, B = cfg
    if maxdiff and max(widths) > 16:
        pytest.skip("the oracle's multiobjective term is a python loop over kernel pairs")
    n_iter = 2
    net, net_c, net_b, scale, shapes = make_net(ctx, D, Nx, Ny, widths, pools, B)
    try:
        P, N = len(widths), 2 * len(widths)
        x = O.synth_frames(11, B, D, Nx, Ny)
        traces = net.fft_step(x, del0=0.2, maxdiff=maxdiff, n_iter=n_iter, fft_l=0)
        trained = [net.get_conv(n) for n in range(N)]
        # reference-shaped path of this engine: layers in real space, then backprop_fft per pair
        layers, _ = ctx.autoenc_fft(x, net_c, net_b, scale, shapes, None, 1)
        layers = [np.ascontiguousarray(t) for t in layers]
        for n in range(P):
            c, f, b, p = (net_c[n].copy(), net_c[N - 1 - n].copy(), net_b[n].copy(), net_b[N - 1 - n].copy())
            li, lo = 2 * n + 1, 2 * N - 1 - 2 * n
            tr = ctx.backprop_fft(layers[li], layers[li], layers[lo], c, f, b, p, 0.2, maxdiff, n_iter)
            assert np.allclose(traces[n], tr, rtol=1e-4), (n, traces[n], tr)
            for got, want, base in ((trained[n][0], c, net_c[n]), (trained[N - 1 - n][0], f, net_c[N - 1 - n]),
                                    (trained[n][1], b, net_b[n]), (trained[N - 1 - n][1], p, net_b[N - 1 - n])):
                assert O.rel_l2(got, want) < 2e-5, n
                assert O.rel_l2(got.astype(np.float64) - base, want.astype(np.float64) - base) < 1e-3, n
        # fp64 oracle, batched: forward per frame, then backprop_fft on the stacked layers
        olayers = [O.autoenc_fft(x[k], net_c, net_b, scale, None, 1)[0] for k in range(B)]
        for n in range(P):
            li, lo = 2 * n + 1, 2 * N - 1 - 2 * n
            inp = np.stack([olayers[k][li] for k in range(B)])
            out = np.stack([olayers[k][lo] for k in range(B)])
            want = O.backprop_fft(inp, inp, out, net_c[n], net_c[N - 1 - n], net_b[n], net_b[N - 1 - n], 0.2, maxdiff, n_iter)
            assert np.allclose(traces[n], want["mse"], rtol=2e-4), (n, traces[n], want["mse"])
            for got, key, base in ((trained[n][0], "c", net_c[n]), (trained[N - 1 - n][0], "f", net_c[N - 1 - n]),
                                   (trained[n][1], "b", net_b[n]), (trained[N - 1 - n][1], "p", net_b[N - 1 - n])):
                assert O.rel_l2(got, want[key]) < 1e-4, (n, key)
                assert O.rel_l2(got.astype(np.float64) - base, want[key] - base) < 2e-3, (n, key)
    finally:
        net.close()


@pytest.mark.parametrize("gram_loop", [True, False])
def test_net_fft_step_uses_tensor_cores_and_no_layer_transforms(ctx, gram_loop, monkeypatch):
    """At the c3 channel widths the forward runs the tcgen05 contraction for pairs 1 and 2 and transforms only the frames
    (one R2C) and the reconstruction (one C2R): no per-layer C2R/R2C round trips.  Training: by default every pair's iteration
    loop runs on the per-bin Gram matrices (spec_gram.cu: one statistics pass over the frames per pair, then n_iter + 1
    kernel-spectrum-sized evaluations); with AEFFT_NO_GRAM_LOOP / AEFFT_NO_GRAM the per-frame forms (tensor-core pairs:
    adjoint + 2 outer products + 2 re-forward contractions per iteration)."""
    if not gram_loop:
        monkeypatch.setenv("AEFFT_NO_GRAM_LOOP", "1")
        monkeypatch.setenv("AEFFT_NO_GRAM", "1")
    net, *_ = make_net(ctx, 3, 64, 64, [16, 32, 64], [2, 2, 2], 16)
    try:
        x = O.synth_frames(3, 16, 3, 64, 64)
        net.fft_step(x, n_iter=1, fft_l=0, want_mse=False)  # plans / allocates
        ctx.profile_enable(True)
        net.fft_step(x, n_iter=1, fft_l=0, want_mse=False)
        rec = {r["name"]: r["launches"] for r in ctx.profile_records()}
        ctx.profile_enable(False)
        if gram_loop:
            assert rec.get("spec_contract_tc", 0) == 4 and rec.get("spec_outer_tc", 0) == 0, rec
            assert rec.get("spec_gram_stats", 0) == 3 and rec.get("spec_gram_iter", 0) == 6, rec
        else:
            assert rec.get("spec_contract_tc", 0) == 10 and rec.get("spec_outer_tc", 0) == 4, rec
            assert "spec_gram_stats" not in rec, rec
        assert rec.get("fft_rows_r2c", 0) == 1 and rec.get("fft_rows_c2r", 0) == 1, rec
    finally:
        net.close()


@pytest.mark.parametrize("size,pool", [((1024, 512), 2), ((256, 2048), 4), ((512, 512), 2)])
def test_net_fft_forward_fused_pooling_at_full_resolution(ctx, size, pool):
    """The frame transform writes the pooled spectrum directly and the reconstruction's inverse transform embeds on the fly
    (fft_kernels.cu: launch_fft_r2c_pooled / launch_fft_c2r_embedded) -- at the BASELINE resolutions (long multi-pass
    transforms), against the reference-shaped path, which transforms at full resolution and pools with the resize kernel."""
    Nx, Ny = size
    net, net_c, net_b, scale, shapes = make_net(ctx, 3, Nx, Ny, [4], [pool], 2)
    try:
        x = O.synth_frames(21, 2, 3, Nx, Ny)
        ctx.profile_enable(True)
        net.fft_forward(x, fft_l=1)
        ctx.sync()
        rec = {r["name"]: r["launches"] for r in ctx.profile_records()}
        ctx.profile_enable(False)
        assert "spec_resize" not in rec, rec          # both poolings were fused into the transforms
        got = [net.layer(l) for l in range(net.num_layers)]
        capi, _ = ctx.autoenc_fft(x, net_c, net_b, scale, shapes, None, 1)
        for l in range(len(shapes)):
            assert O.rel_l2(got[l], capi[l]) < 2e-5, l
        want, _ = O.autoenc_fft(x[0], net_c, net_b, scale, None, 0)
        assert O.rel_l2(got[-1][0], want[-1]) < 2e-5
    finally:
        net.close()


@pytest.mark.parametrize("cfg", [(3, 64, 32, [4, 5], [2, 2], 3), (3, 64, 64, [16, 32], [2, 2], 16), (1, 32, 64, [8, 16], [1, 2], 16),
                                 (3, 64, 128, [8, 16, 32], [2, 2, 2], 16), (3, 80, 120, [8, 16], [2, 2], 16)])
def test_net_fft_forward_fused_with_level_changes_equals_unfused(ctx, cfg, monkeypatch):
    """fft_l <= 0: the image-side convs run fused with the spectral pooling after / the up-sampling before them and compute
    only the kept bins (net_fft.cu: conv_then_pool / unpool_then_conv); a tensor-core level followed by a pooling does the
    same when its training runs on the Gram matrices (conv_then_pool_tc; the last configuration).  Same arithmetic per kept
    bin, so the reconstruction and the trained kernels equal those of the unfused path (AEFFT_NO_FWD_FUSE) bit for bit."""
    D, Nx, Ny, widths, pools, B = cfg
    out = []
    # fused + decoder on the support grid (default) | fused, dense decoder | nothing fused
    for mode in ("default", "dense_decoder", "unfused"):
        nofuse = mode == "unfused"
        monkeypatch.delenv("AEFFT_NO_FWD_FUSE", raising=False)
        monkeypatch.delenv("AEFFT_NO_SPARSE_DECODER", raising=False)
        if mode == "unfused":
            monkeypatch.setenv("AEFFT_NO_FWD_FUSE", "1")
        if mode == "dense_decoder":
            monkeypatch.setenv("AEFFT_NO_SPARSE_DECODER", "1")
        net, *_ = make_net(ctx, D, Nx, Ny, widths, pools, B)
        try:
            x = O.synth_frames(5, B, D, Nx, Ny)
            ctx.profile_enable(True)
            traces = net.fft_step(x, del0=0.2, maxdiff=0, n_iter=2, fft_l=0)
            names = {r["name"] for r in ctx.profile_records()}
            ctx.profile_enable(False)
            assert ("spec_contract_reg_pool" in names) == (not nofuse), names
            # (the decoder runs on the support grid only when the reconstruction's embedding inverse transform exists:
            # power-of-two frames)
            sparse = mode == "default" and all(v & (v - 1) == 0 for v in (Nx, Ny))
            assert ("spec_contract_reg_embed" in names) == (not nofuse and not sparse), names
            assert ("spec_contract_reg_support" in names) == sparse, names
            out.append((np.array(traces), net.layer(net.num_layers - 1).copy(),
                        [net.get_conv(n)[0].copy() for n in range(2 * net.num_pairs)]))
        finally:
            net.close()
    for other in out[1:]:
        assert np.array_equal(out[0][0], other[0])
        assert np.array_equal(out[0][1], other[1])
        for a, b in zip(out[0][2], other[2]):
            assert np.array_equal(a, b)


def test_net_fft_step_at_the_camera_format(ctx):
    """640 x 480 frames in momentum space (lengths 640 = 2^7 5 and 480 = 2^5 3 5 and their pooled levels run on
    the mixed-radix transforms): one step of the c3 stack against the reference-shaped C-ABI path of this engine (per-layer
    transforms), which the small mixed-radix cases above pin against the fp64 oracle."""
    net, net_c, net_b, scale, shapes = make_net(ctx, 3, 640, 480, [16, 32, 64], [2, 2, 2], 16)
    try:
        P, N = 3, 6
        x = O.synth_frames(13, 16, 3, 640, 480)
        traces = net.fft_step(x, del0=0.2, maxdiff=0, n_iter=1, fft_l=0)
        trained = [net.get_conv(n) for n in range(N)]
        layers, _ = ctx.autoenc_fft(x, net_c, net_b, scale, shapes, None, 1)
        layers = [np.ascontiguousarray(t) for t in layers]
        assert O.rel_l2(net.layer(len(shapes) - 1), layers[-1]) < 2e-5
        for n in range(P):
            c, f, b, p = (net_c[n].copy(), net_c[N - 1 - n].copy(), net_b[n].copy(), net_b[N - 1 - n].copy())
            li, lo = 2 * n + 1, 2 * N - 1 - 2 * n
            tr = ctx.backprop_fft(layers[li], layers[li], layers[lo], c, f, b, p, 0.2, 0, 1)
            assert np.allclose(traces[n], tr, rtol=1e-4), (n, traces[n], tr)
            for got, want in ((trained[n][0], c), (trained[N - 1 - n][0], f), (trained[n][1], b), (trained[N - 1 - n][1], p)):
                assert O.rel_l2(got, want) < 2e-5, n
    finally:
        net.close()
