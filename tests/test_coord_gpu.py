"""GPU parity tests of the coordinate-space path, through the C ABI (libaefft.so), against
  (1) the numpy oracle (oracle/oracle_np.py, pinned on CPU by tests/test_oracle_cpu.py),
  (2) the UNMODIFIED reference run live on the same GPU (oracle/_ref/libref.so) when it travelled with the repo,
  (3) committed golden vectors of the reference's CUDA path (tests/golden/gpu_golden.npz) when present.
Tolerances (north_star): 1e-4 relative L2 on weights / reconstructions (fp32 mode); updates (delta w) at 1e-3.
Integer-valued work (Pool, Portion, synthetic frames) is bit-exact."""
import os

import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL_W = 1e-4
TOL_DW = 1e-3


def make_case(seed, dM, dD, Nk, Nl, Nx, Ny, B=None, wscale=0.2, conv="cuda"):
    rng = np.random.default_rng(seed)
    lead = () if B is None else (B,)
    inp = np.floor(rng.random(lead + (dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * wscale).astype(np.float32)
    f = ((rng.random((dD, dM, Nk, Nl)) * 2 - 1) * wscale).astype(np.float32)
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    p = (rng.random(dD) * 2 - 1).astype(np.float32)
    fwd = O.conv_gpu if conv == "cuda" else O.conv_cpu
    hin = fwd(inp, c, b).astype(np.float32)
    out = fwd(hin, f, p).astype(np.float32)
    return dict(inp=inp, hin=hin, out=out, c=c, b=b, f=f, p=p)


def zeros_state(cs):
    st = {}
    for k in ("dc", "ddc"):
        st[k] = np.zeros_like(cs["c"])
    for k in ("df", "ddf"):
        st[k] = np.zeros_like(cs["f"])
    for k in ("db", "ddb"):
        st[k] = np.zeros_like(cs["b"])
    for k in ("dp", "ddp"):
        st[k] = np.zeros_like(cs["p"])
    return st


# ------------------------------------------------------------------------------------------------ bit-exact pieces
def test_synth_frames_bit_exact(ctx):
    got = ctx.synth_frames(1234, 3, 3, 40, 24, b0=5)
    assert np.array_equal(got, O.synth_frames(1234, 3, 3, 40, 24, b0=5))


@pytest.mark.parametrize("scale,shape,oshape", [(2, (3, 12, 10), (6, 5)), (1, (2, 9, 7), (9, 7)), (3, (1, 12, 9), (4, 3)),
                                               (-2, (3, 6, 5), (12, 10)), (-3, (2, 4, 3), (12, 9)),
                                               # the vectorised factor-2 fast paths (row length % 8 / % 4)
                                               (2, (3, 12, 16), (6, 8)), (2, (5, 20, 40), (10, 20)), (-2, (3, 6, 8), (12, 16)),
                                               (-2, (4, 10, 20), (20, 40))])
def test_pool_bit_exact(ctx, scale, shape, oshape):
    rng = np.random.default_rng(1)
    x = (rng.random((2,) + shape) * 300 - 40).astype(np.float32)  # negatives and fractions exercise the int quirk (N1)
    assert np.array_equal(ctx.pool(x, scale, oshape), O.pool(x, scale, oshape))


def test_portion_bit_exact(ctx):
    rng = np.random.default_rng(2)
    x = rng.random((2, 3, 20, 14)).astype(np.float32)
    for q in (1, 2, 3):
        want = O.portion(x, x, x, q)[0]
        assert np.array_equal(ctx.portion(x, q), want)


# ------------------------------------------------------------------------------------------------ forward conv
CONV_SHAPES = [(4, 3, 5, 5, 20, 14), (16, 3, 5, 5, 70, 66), (8, 1, 5, 5, 64, 48), (3, 2, 3, 3, 9, 13), (2, 1, 7, 7, 16, 16),
               (32, 16, 5, 5, 33, 65), (5, 4, 5, 3, 17, 19), (20, 6, 5, 5, 31, 130)]


@pytest.mark.parametrize("dims", CONV_SHAPES)
@pytest.mark.parametrize("conv", ["cuda", "cpu"])
def test_conv_fwd_vs_oracle(ctx, dims, conv):
    cs = make_case(5, *dims, conv=conv)
    got = ctx.conv_fwd(cs["inp"], cs["c"], cs["b"], A.CONV_CUDA if conv == "cuda" else A.CONV_CPU)
    want = (O.conv_gpu if conv == "cuda" else O.conv_cpu)(cs["inp"], cs["c"], cs["b"])
    assert O.rel_l2(got, want) < 1e-5


def test_conv_fwd_batched_equals_per_frame(ctx):
    cs = make_case(6, 6, 3, 5, 5, 24, 40, B=3)
    got = ctx.conv_fwd(cs["inp"], cs["c"], cs["b"])
    for n in range(3):
        assert O.rel_l2(got[n], O.conv_gpu(cs["inp"][n], cs["c"], cs["b"])) < 1e-5


def test_conv_fwd_vs_live_reference(ctx, ref):
    if ref is None:
        pytest.skip("libref.so not present")
    for dims in [(4, 3, 5, 5, 20, 14), (3, 2, 3, 3, 9, 13), (2, 1, 7, 7, 16, 16), (16, 3, 5, 5, 64, 48)]:
        cs = make_case(7, *dims)
        want = ref.conv_gpu(cs["inp"], cs["c"], cs["b"])
        got = ctx.conv_fwd(cs["inp"], cs["c"], cs["b"])
        assert O.rel_l2(got, want) < 1e-5, dims
        want_cpu = ref.conv_cpu(cs["inp"], cs["c"], cs["b"])
        got_cpu = ctx.conv_fwd(cs["inp"], cs["c"], cs["b"], A.CONV_CPU)
        assert O.rel_l2(got_cpu, want_cpu) < 1e-5, dims


# ------------------------------------------------------------------------------------------------ training step
def run_product(ctx, mode, cs, st, steps=1, quirks=A.QUIRKS_ALL, delmax=0.2, alpha=0.9):
    w = {k: cs[k].copy() for k in "cbfp"}
    s = {k: v.copy() for k, v in st.items()}
    mses = []
    for _ in range(steps):
        mses.append(ctx.backprop_coord(mode, cs["inp"], cs["out"], cs["hin"], w["c"], w["b"], w["f"], w["p"], s["dc"],
                                       s["db"], s["df"], s["dp"], s["ddc"], s["ddb"], s["ddf"], s["ddp"], delmax, alpha,
                                       1, quirks))
    w.update(s)
    w["mse"] = mses
    return w


def check_weights(got, want, base, keys="cbfp", tol_w=TOL_W, tol_dw=TOL_DW):
    for k in keys:
        assert O.rel_l2(got[k], want[k]) < tol_w, f"{k}: {O.rel_l2(got[k], want[k])}"
        dw_want = np.asarray(want[k], np.float64) - base[k]
        if np.linalg.norm(dw_want) > 0:
            r = O.rel_l2(np.asarray(got[k], np.float64) - base[k], dw_want)
            assert r < tol_dw, f"delta {k}: {r}"


SQUARE = [(4, 3, 5, 5, 16, 16), (3, 2, 3, 3, 12, 12), (16, 3, 5, 5, 48, 48), (2, 1, 7, 7, 20, 20)]
RECT = [(4, 3, 5, 5, 20, 14), (16, 3, 5, 5, 40, 72), (8, 5, 3, 3, 33, 17)]


@pytest.mark.parametrize("dims", SQUARE + RECT)
def test_backprop_sym_vs_oracle(ctx, dims):
    cs = make_case(8, *dims)
    cs["f"] = np.ascontiguousarray(np.swapaxes(cs["c"], 0, 1))
    cs["out"] = O.conv_gpu(cs["hin"], cs["f"], cs["p"]).astype(np.float32)
    st = zeros_state(cs)
    got = run_product(ctx, A.MODE_CUDA_REF_SYM, cs, st)
    want = O.backprop_gpu_cc(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], **st, delmax=0.2,
                             alpha=0.9)
    check_weights(got, want, cs)
    assert abs(got["mse"][0] - want["mse"]) <= 1e-4 * abs(want["mse"])
    for k in ("dc", "db", "dp", "ddc", "ddb", "ddp"):
        assert O.rel_l2(got[k], want[k]) < TOL_DW, k
    assert np.array_equal(got["f"], np.swapaxes(got["c"], 0, 1))  # tied weights stay tied (backproplib.cu:622)


@pytest.mark.parametrize("dims", SQUARE)
@pytest.mark.parametrize("quirks", [A.QUIRKS_ALL, 0])
def test_backprop_cuda_ref_vs_oracle(ctx, dims, quirks):
    cs = make_case(9, *dims)
    st = zeros_state(cs)
    got = run_product(ctx, A.MODE_CUDA_REF, cs, st, quirks=quirks)
    want = O.backprop_gpu(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], **st, delmax=0.2,
                          alpha=0.9, quirks=bool(quirks))
    check_weights(got, want, cs)
    for k in ("ddc", "ddf", "ddb", "ddp"):
        assert O.rel_l2(got[k], want[k]) < TOL_DW, k


@pytest.mark.parametrize("dims", [(8, 1, 5, 5, 40, 30), (4, 3, 5, 5, 14, 12), (3, 2, 3, 3, 10, 10), (2, 2, 7, 7, 18, 21)])
def test_backprop_cpu_ref_vs_oracle(ctx, dims):
    cs = make_case(10, *dims, conv="cpu")
    got = run_product(ctx, A.MODE_CPU_REF, cs, zeros_state(cs), delmax=0.5)
    want = O.backprop_cpu(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], 0.5)
    check_weights(got, want, cs)
    assert abs(got["mse"][0] - want["mse"]) <= 1e-4 * abs(want["mse"])


def test_backprop_cpu_ref_golden(ctx):
    """The committed golden vectors of the compiled netlib.cpp backprop() (incl. the config-1 shape class)."""
    G = np.load(os.path.join(GOLDEN, "cpu_golden.npz"))
    for tag in ("c1", "d3", "k3"):
        cs = {k: np.ascontiguousarray(G[f"bp_{tag}_{k}"]) for k in "inp out hin c b f p".split()}
        got = run_product(ctx, A.MODE_CPU_REF, cs, zeros_state(cs), delmax=float(G[f"bp_{tag}_delta"]))
        want = {k: G[f"bp_{tag}_new_{k}"] for k in "cbfp"}
        check_weights(got, want, cs)


@pytest.mark.parametrize("mode", ["sym", "cuda", "cpu"])
def test_backprop_batched_mean_gradient(ctx, mode):
    dims = (6, 3, 5, 5, 24, 24)
    cs = make_case(11, *dims, B=3, conv="cpu" if mode == "cpu" else "cuda")
    st = zeros_state(cs)
    if mode == "sym":
        cs["f"] = np.ascontiguousarray(np.swapaxes(cs["c"], 0, 1))
        got = run_product(ctx, A.MODE_CUDA_REF_SYM, cs, st)
        want = O.backprop_gpu_cc(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], **st, delmax=0.2, alpha=0.9)
    elif mode == "cuda":
        got = run_product(ctx, A.MODE_CUDA_REF, cs, st)
        want = O.backprop_gpu(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], **st, delmax=0.2, alpha=0.9)
    else:
        got = run_product(ctx, A.MODE_CPU_REF, cs, st)
        want = O.backprop_cpu(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], 0.2)
    check_weights(got, want, cs)


def test_backprop_vs_live_reference(ctx, ref):
    """Two consecutive steps (momentum carried) against the unmodified backprop_gpu / backprop_gpu_cc / backprop."""
    if ref is None:
        pytest.skip("libref.so not present")
    for dims in [(4, 3, 5, 5, 16, 16), (3, 2, 3, 3, 12, 12), (8, 3, 5, 5, 32, 32)]:
        for sym in (0, 1):
            cs = make_case(12, *dims)
            if sym:
                cs["f"] = np.ascontiguousarray(np.swapaxes(cs["c"], 0, 1))
            cs["hin"] = ref.conv_gpu(cs["inp"], cs["c"], cs["b"])
            cs["out"] = ref.conv_gpu(cs["hin"], cs["f"], cs["p"])
            st = zeros_state(cs)
            names = "c b f p dc db df dp ddc ddb ddf ddp".split()
            # tied weights: two consecutive steps (momentum carried).  Independent weights: ONE step, because
            # backprop_gpu's dF term (quirk C3) reads hin_flat[m*P + (i-ik)*Nx + (j-ik)], which for m == 0 falls below
            # the buffer and for m == dM-1 beyond its end (undefined behaviour: the compiled reference picks up
            # whatever the neighbouring cudaMalloc holds).  Those reads are defined as 0 here, so channels
            # 0 and dM-1 of f are excluded; after a second step the garbage would also have leaked into c through dh.
            steps = 2 if sym else 1
            want = dict(cs, **st)
            for _ in range(steps):
                want = ref.backprop_gpu(sym, cs["inp"], cs["out"], cs["hin"], *[want[k] for k in names], 0.2, 0.9, 1)
            got = run_product(ctx, A.MODE_CUDA_REF_SYM if sym else A.MODE_CUDA_REF, cs, st, steps=steps)
            if sym:
                check_weights(got, want, cs)
                continue
            check_weights(got, want, cs, keys="cbp")
            inner = slice(1, dims[0] - 1)
            fg, fw, f0 = (np.asarray(t, np.float64)[:, inner] for t in (got["f"], want["f"], cs["f"]))
            assert O.rel_l2(fg, fw) < TOL_W
            assert O.rel_l2(fg - f0, fw - f0) < TOL_DW
    cs = make_case(13, 8, 1, 5, 5, 40, 30, conv="cpu")
    want = ref.backprop_cpu(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], 0.2)
    got = run_product(ctx, A.MODE_CPU_REF, cs, zeros_state(cs))
    check_weights(got, want, cs)


def test_gpu_golden_coordinate(ctx):
    path = os.path.join(GOLDEN, "gpu_golden.npz")
    if not os.path.exists(path):
        pytest.skip("gpu_golden.npz not generated yet (tests/golden/make_golden_gpu.py under gpurun)")
    G = np.load(path)
    for tag in "abc":
        got = ctx.conv_fwd(np.ascontiguousarray(G[f"convg_{tag}_x"]), np.ascontiguousarray(G[f"convg_{tag}_c"]),
                           np.ascontiguousarray(G[f"convg_{tag}_b"]))
        assert O.rel_l2(got, G[f"convg_{tag}_out"]) < 1e-5
    names = "c b f p dc db df dp ddc ddb ddf ddp".split()
    for tag in ("s5", "s3"):
        for sym in (0, 1):
            key = f"bpg_{tag}_{sym}"
            cs = {k: np.ascontiguousarray(G[f"{key}_{k}"]) for k in ["inp", "hin", "out"] + names}
            st = {k: cs[k] for k in names[4:]}
            steps = 2 if sym else 1  # see test_backprop_vs_live_reference for why independent weights stop at one step
            got = run_product(ctx, A.MODE_CUDA_REF_SYM if sym else A.MODE_CUDA_REF, cs, st, steps=steps)
            want = {k: G[f"{key}_step{steps}_{k}"] for k in names}
            if sym:
                check_weights(got, want, cs)
            else:
                check_weights(got, want, cs, keys="cbp")
                dM = cs["c"].shape[0]
                fg, fw, f0 = (np.asarray(t, np.float64)[:, 1:dM - 1] for t in (got["f"], want["f"], cs["f"]))
                assert O.rel_l2(fg, fw) < TOL_W and O.rel_l2(fg - f0, fw - f0) < TOL_DW


# ------------------------------------------------------------------------------------------------ device-resident net
def oracle_forward(x, net_c, net_b, scale):
    layers = [x]
    N = len(net_c)
    for n in range(N):
        if n < N // 2:
            pin = O.pool(layers[-1], scale[n])
            layers.append(pin)
            layers.append(O.conv_gpu(pin, net_c[n], net_b[n]).astype(np.float32))
        else:
            h = O.conv_gpu(layers[-1], net_c[n], net_b[n]).astype(np.float32)
            layers.append(h)
            D, Nx, Ny = h.shape[-3:]
            s = -scale[n]
            layers.append(O.pool(h, scale[n], (Nx * s, Ny * s)))
    return layers


def test_net_step_matches_composed_oracle(ctx):
    import ctypes

    ctypes.CDLL("libc.so.6").srand(1234)
    B, D, Nx, Ny = 2, 3, 32, 24
    net = A.Net(ctx, D, Nx, Ny, B)
    net.add_layer(4, 1, 1, 2, 0.3)
    net.add_layer(6, 1, 1, 2, 0.3)
    assert net.num_pairs == 2 and net.num_layers == 9
    rng = O.GlibcRand(1234)
    c0, b0 = O.init_conv(rng, 4, 3, 5, 5, 0.3)
    f0, p0 = O.init_conv(rng, 3, 4, 5, 5, 0.3)
    for shp in ((4, 3), (3, 4), (4, 3), (3, 4)):  # Init_conv(dc/df/ddc/ddf, ..., 0) still draws (autoencoder.cpp:103-107)
        O.init_conv(rng, shp[0], shp[1], 5, 5, 0.0)
    c1, b1 = O.init_conv(rng, 6, 4, 5, 5, 0.3)
    f1, p1 = O.init_conv(rng, 4, 6, 5, 5, 0.3)
    net_c, net_b, scale = [c0, c1, f1, f0], [b0, b1, p1, p0], [2, 2, -2, -2]
    for n in range(4):
        gc, gb = net.get_conv(n)
        assert np.array_equal(gc, net_c[n]) and np.array_equal(gb, net_b[n])  # seeded weights are bit-identical
        assert net.conv_dims(n)[4] == scale[n]
    x = O.synth_frames(1234, B, D, Nx, Ny)
    mse = np.zeros(2, np.float32)
    net.step(x, A.MODE_CUDA_REF_SYM, mse=mse)
    layers = oracle_forward(x, net_c, net_b, scale)
    for l in range(9):
        assert O.rel_l2(net.layer(l), layers[l]) < 1e-5, l
    # pair 0: in=L1 hin=L2 out=L7 ; pair 1: in=L3 hin=L4 out=L5  (autoencoder.cpp:161-169)
    for n_l, (i, h, o, c, b, f, p) in enumerate([(1, 2, 7, c0, b0, f0, p0), (3, 4, 5, c1, b1, f1, p1)]):
        z = lambda a: np.zeros_like(a)
        want = O.backprop_gpu_cc(layers[i], layers[o], layers[h], c, b, f, p, z(c), z(b), z(f), z(p), z(c), z(b), z(f),
                                 z(p), 0.2, 0.9)
        gc, gb = net.get_conv(n_l)
        gf, gp = net.get_conv(3 - n_l)
        got = dict(c=gc, b=gb, f=gf, p=gp)
        check_weights(got, want, dict(c=c, b=b, f=f, p=p))
        assert abs(mse[n_l] - want["mse"]) <= 1e-4 * abs(want["mse"])
    net.delete_layer()
    assert net.num_pairs == 1 and net.num_layers == 5
    with pytest.raises(A.AefftError):
        net.delete_layer()  # never the last remaining pair (autoencoder.cpp:434)
    net.close()


def test_data_parallel_split_equals_full_batch(ctx):
    """gradients(frames 0..1) + gradients(frames 2..3) -> sum -> update == one 4-frame step (what the all-reduce does)."""
    dims = (6, 3, 5, 5, 24, 24)
    cs = make_case(14, *dims, B=4)
    cs["f"] = np.ascontiguousarray(np.swapaxes(cs["c"], 0, 1))
    st = zeros_state(cs)
    full = run_product(ctx, A.MODE_CUDA_REF_SYM, cs, st)
    dM, dD, Nk, Nl, Nx, Ny = dims
    n = int(A.lib().aefft_coord_gbuf_len(A.MODE_CUDA_REF_SYM, dD, dM, Nk, Nl))
    dev = {k: ctx.to_device(cs[k]) for k in "inp out hin c b f p".split()}
    dst = {k: ctx.to_device(v) for k, v in st.items()}
    halves = []
    for h in range(2):
        sl = slice(2 * h, 2 * h + 2)
        part = {k: ctx.to_device(cs[k][sl]) for k in ("inp", "out", "hin")}
        g = A.DevBuf(ctx, (n,))
        ctx.coord_gradients(A.MODE_CUDA_REF_SYM, A.QUIRKS_ALL, 2, dD, dM, Nx, Ny, Nk, Nl, part["inp"], part["out"],
                            part["hin"], dev["c"], dev["f"], g)
        halves.append(g.numpy())
    gsum = ctx.to_device(halves[0] + halves[1])
    ctx.coord_update(A.MODE_CUDA_REF_SYM, 4, dD, dM, Nx, Ny, Nk, Nl, gsum, dev["c"], dev["b"], dev["f"], dev["p"],
                     dst["dc"], dst["db"], dst["df"], dst["dp"], dst["ddc"], dst["ddb"], dst["ddf"], dst["ddp"], 0.2, 0.9)
    ctx.sync()
    for k in "cbfp":
        assert O.rel_l2(dev[k].numpy(), full[k]) < 1e-6, k


def test_image_to_spin_u8_bit_exact(ctx):
    """aefft_net_set_frames_u8 = ImageToSpin_C (netlib.cpp:37-50): spin[d][i][j] = (float)img(row j, col i)[d]."""
    B, D, Nx, Ny = 3, 3, 70, 44
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(B, Ny, Nx, D), dtype=np.uint8)
    net = A.Net(ctx, D, Nx, Ny, B)
    try:
        net.add_layer(4, 1, 1, 2, 1.0)
        net.set_frames_u8(img)
        ctx.sync()
        _, _, _, ptr = net.layer_info(0)
        got = np.empty((B, D, Nx, Ny), np.float32)
        ctx.memcpy(got.ctypes.data, ptr, got.nbytes, 1)
    finally:
        net.close()
    want = img.transpose(0, 3, 2, 1).astype(np.float32)
    assert np.array_equal(got, want)


def test_spin_to_image_u8_bit_exact(ctx):
    """aefft_net_get_layer_u8 = SpinToImage_C (netlib.cpp:52-76: clamp(round(v), 0, 255), round half away from zero) and
    SpinToImage_V (:79-92: (uchar)(int)v): the display side of the reference's webcam loop."""
    B, D, Nx, Ny = 2, 3, 38, 52
    rng = np.random.default_rng(9)
    net = A.Net(ctx, D, Nx, Ny, B)
    try:
        net.add_layer(4, 1, 1, 2, 1.0)
        v = (rng.random((B, D, Nx, Ny)) * 300 - 20).astype(np.float32)
        v[0, 0, :4, 0] = [0.5, 1.5, 2.5, -0.5]  # halves: C round() goes away from zero
        _, _, _, ptr = net.layer_info(0)
        ctx.memcpy(ptr, v.ctypes.data, v.nbytes, 0)
        got_c, got_v = net.get_layer_u8(0, 0), net.get_layer_u8(0, 1)
    finally:
        net.close()
    r = np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), np.ceil(v.astype(np.float64) - 0.5))
    want_c = np.clip(r, 0, 255).astype(np.uint8).transpose(0, 3, 2, 1)
    want_v = (np.trunc(v).astype(np.int64) & 255).astype(np.uint8).transpose(0, 3, 2, 1)
    assert np.array_equal(got_c, want_c)
    assert np.array_equal(got_v, want_v)


def test_net_step_cuda_ref_quirks_two_pairs(ctx):
    """aefft_net_step in CUDA_REF mode with all quirks on, two square pairs of DIFFERENT resolution trained back to back
    without a host sync in between (the stale-border kernel of pair 0 must not see pair 1's border geometry), against
    the oracle's backprop_gpu per pair."""
    import ctypes

    ctypes.CDLL("libc.so.6").srand(77)
    B, D, Nx, Ny = 2, 3, 32, 32
    net = A.Net(ctx, D, Nx, Ny, B)
    try:
        net.add_layer(5, 1, 1, 2, 0.3)
        net.add_layer(4, 1, 1, 2, 0.3)
        convs = [net.get_conv(n) for n in range(4)]
        net_c, net_b, scale = [c for c, _ in convs], [b for _, b in convs], [2, 2, -2, -2]
        x = O.synth_frames(77, B, D, Nx, Ny)
        net.step(x, A.MODE_CUDA_REF, quirks=A.QUIRKS_ALL)  # no mse pointer: nothing synchronises between the pairs
        ctx.sync()
        layers = oracle_forward(x, net_c, net_b, scale)
        for n_l, (i, h, o) in enumerate([(1, 2, 7), (3, 4, 5)]):
            c, b = convs[n_l]
            f, p = convs[3 - n_l]
            z = lambda a: np.zeros_like(a)
            want = O.backprop_gpu(layers[i], layers[o], layers[h], c, b, f, p, z(c), z(b), z(f), z(p), z(c), z(b), z(f),
                                  z(p), 0.2, 0.9, quirks=True)
            gc, gb = net.get_conv(n_l)
            gf, gp = net.get_conv(3 - n_l)
            check_weights(dict(c=gc, b=gb, f=gf, p=gp), want, dict(c=c, b=b, f=f, p=p))
    finally:
        net.close()


def test_ten_steps_bf16x3_vs_oracle(ctx):
    """north_star gate: updated weights and reconstructions after N steps.  10 steps of the symmetric-weights net step on
    a 3-pair stack (3->8->16->32, the c2 shape class at reduced size) in the DEFAULT tensor-core precision (BF16X3),
    seeded Init_conv weights (rmax 0.3), rate 0.05 (at the app's 0.2 this stack diverges within 8 steps, in
    the oracle as well), synthetic frames; the oracle replays forward + backprop_gpu_cc per pair in fp64."""
    import ctypes

    ctypes.CDLL("libc.so.6").srand(1234)
    B, D, Nx, Ny, widths, steps = 4, 3, 64, 48, [8, 16, 32], 10
    prec = ctx.precision
    ctx.set_precision(A.PRECISION_BF16X3)
    net = A.Net(ctx, D, Nx, Ny, B)
    try:
        for m in widths:
            net.add_layer(m, 1, 1, 2, 0.3)
        P = len(widths)
        for n in range(P):
            net.set_symmetric(n)
        convs = [net.get_conv(n) for n in range(2 * P)]
        start_c = [c.copy() for c, _ in convs]
        net_c = [c.astype(np.float64) for c, _ in convs]
        net_b = [b.astype(np.float64) for _, b in convs]
        scale = [2] * P + [-2] * P
        mom = [dict(dc=np.zeros_like(net_c[n]), db=np.zeros_like(net_b[n]), df=np.zeros_like(net_c[2 * P - 1 - n]),
                    dp=np.zeros_like(net_b[2 * P - 1 - n])) for n in range(P)]
        ctx.profile_enable(True)
        for s in range(steps):
            x = O.synth_frames(1234, B, D, Nx, Ny, b0=s * B)  # new frames every step
            net.step(x, A.MODE_CUDA_REF_SYM, delmax=0.05, alpha=0.9)
            layers = oracle_forward(x, net_c, net_b, scale)
            for n in range(P):
                i, h, o = 2 * n + 1, 2 * n + 2, len(layers) - 2 - 2 * n
                st = mom[n]
                zc, zb, zf, zp = (np.zeros_like(t) for t in (st["dc"], st["db"], st["df"], st["dp"]))
                w = O.backprop_gpu_cc(layers[i], layers[o], layers[h], net_c[n], net_b[n], net_c[2 * P - 1 - n],
                                      net_b[2 * P - 1 - n], st["dc"], st["db"], st["df"], st["dp"], zc, zb, zf, zp, 0.05, 0.9)
                net_c[n], net_b[n], net_c[2 * P - 1 - n], net_b[2 * P - 1 - n] = w["c"], w["b"], w["f"], w["p"]
                mom[n] = dict(dc=w["dc"], db=w["db"], df=w["df"], dp=w["dp"])
        names = {r["name"] for r in ctx.profile_records()}
        ctx.profile_enable(False)
        assert "wgrad_ts" in names or any(n.startswith("wgrad_t") for n in names), names
        ctx.sync()
        for n in range(2 * P):
            gc, gb = net.get_conv(n)
            assert O.rel_l2(gc, net_c[n]) < 1e-4, ("c", n, O.rel_l2(gc, net_c[n]))
            assert O.rel_l2(gb, net_b[n]) < 1e-4, ("b", n, O.rel_l2(gb, net_b[n]))
            # accumulated update over the 10 steps (weights), 1e-3
            assert O.rel_l2(gc.astype(np.float64) - start_c[n], net_c[n] - start_c[n]) < 1e-3, ("dc", n)
        # reconstructions of a fresh batch with the trained weights
        x = O.synth_frames(1234, B, D, Nx, Ny, b0=steps * B)
        net.forward(x)
        want = oracle_forward(x, net_c, net_b, scale)
        assert O.rel_l2(net.layer(2 * 2 * P), want[-1]) < 1e-4
    finally:
        ctx.set_precision(prec)
        net.close()


@pytest.mark.parametrize("dims", [(16, 3, 5, 5, 240, 240), (32, 16, 5, 5, 120, 120), (64, 32, 5, 5, 60, 60)])
def test_config2_square_twin_vs_live_reference(ctx, ref, dims):
    """the compiled backprop_gpu_cc is only defined on square frames (quirk C2), so every pair of
    config 2 gets a square twin with its channel widths (16x3 @ 240^2, 32x16 @ 120^2, 64x32 @ 60^2, one frame) and the
    streaming tensor-core kernels (default BF16X3 precision) are compared with the UNMODIFIED reference run live."""
    if ref is None:
        pytest.skip("libref.so not present")
    dM, dD, Nk, Nl, Nx, Ny = dims
    rng = np.random.default_rng(dM)
    inp = np.floor(rng.random((dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * (1.0 / dD)).astype(np.float32)
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    p = (rng.random(dD) * 2 - 1).astype(np.float32)
    f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
    hin = ref.conv_gpu(inp, c, b)
    out = ref.conv_gpu(hin, f, p)
    cs = dict(inp=inp, hin=hin, out=out, c=c, b=b, f=f, p=p)
    st = zeros_state(cs)
    names = "c b f p dc db df dp ddc ddb ddf ddp".split()
    want = ref.backprop_gpu(1, inp, out, hin, *[dict(cs, **st)[k] for k in names], 0.2, 0.9, 1)
    prec = ctx.precision
    ctx.set_precision(A.PRECISION_BF16X3)
    try:
        assert O.rel_l2(ctx.conv_fwd(inp, c, b), hin) < 1e-4
        ctx.profile_enable(True)
        got = run_product(ctx, A.MODE_CUDA_REF_SYM, cs, st)
        kernels = {r["name"] for r in ctx.profile_records()}
        ctx.profile_enable(False)
    finally:
        ctx.set_precision(prec)
    assert any(k.startswith("wgrad_t") for k in kernels), kernels  # a tensor-core weight-gradient kernel ran
    check_weights(got, want, cs)
