"""The headless replay driver (autoencoder-fft_b200/tools/aefft_replay.cpp, built by the Makefile) walks the reference
application's key-driven state machine -- add / delete layer from New_Layer_Param.txt, layer cycling with momentum
restart, weight re-draw, symmetric toggle, save / load -- over the C ABI.  The same event sequence is replayed here
through ctypes on the same library; seeded weights, synthetic frames and kernels are deterministic, so every "mse" line
must agree to the printed precision.  Further down the driver's output is checked against the ORACLE (coordinate mode:
backprop_gpu with its quirks; FFT mode, keys f / g: autoenc_fft + backprop_fft), and two C++-only data-parallel ranks
against one rank holding all frames."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O
from conftest import ROOT

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "autoencoder-fft_b200")
PARAM = "Layer_depth 6\nKernel_L_x 1\nKernel_L_y 1\nPooling_scale 2\nMax_Rand_Init 0.3\n"
SCRIPT = "n t3 n t2 i x t2 z p t2 s e t1 l t1 p d t2"
B, D, NX, NY = 2, 3, 48, 48  # square: the driver then reproduces the reference's backprop_gpu defects (quirks) as well


def replica(ctx, tmp):
    """The driver's loop in Python (same ABI calls in the same order)."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1234)
    net = A.Net(ctx, D, NX, NY, B)
    dM, Lk, Ll, scal, rmax = A.load_param(tmp / "New_Layer_Param.txt")
    n_l, sym, frame0, out = 0, 0, 0, []
    for t in SCRIPT.split():
        pairs = net.num_pairs
        N = 2 * pairs - 1
        if t == "n":
            net.add_layer(dM, Lk, Ll, scal, rmax)
            n_l = net.num_pairs - 1
        elif t == "d":
            if pairs > 1:
                net.delete_layer()
                n_l = 0
                net.reset_momentum(n_l)
        elif t in "zx":
            n_l = (n_l + 1) % pairs if t == "z" else (n_l - 1 + pairs) % pairs
            net.reset_momentum(n_l)
        elif t == "e":
            m, d, Nk, Nl, _ = net.conv_dims(n_l)
            c, b = A.init_conv(m, d, Nk, Nl, rmax)
            f, p = A.init_conv(d, m, Nk, Nl, rmax)
            net.set_conv(n_l, c, b)
            net.set_conv(N - n_l, f, p)
        elif t == "p":
            sym = (sym + 1) % 2
            if sym:
                net.set_symmetric(n_l)
        elif t in "sl":
            for io in range(2):
                n = N - n_l if io else n_l
                m, d, Nk, Nl, sc = net.conv_dims(n)
                if t == "s":
                    c, b = net.get_conv(n)
                    A.saveload_conv(tmp / "weights", c, b, sc, n_l, io, True)
                else:
                    c, b = np.empty((m, d, Nk, Nl), np.float32), np.empty(m, np.float32)
                    A.saveload_conv(tmp / "weights", c, b, sc, n_l, io, False)
                    net.set_conv(n, c, b)
        elif t[0] == "t":
            _, _, _, ptr = net.layer_info(0)
            for _ in range(int(t[1:])):
                ctx.synth_frames(1234, B, D, NX, NY, b0=frame0, out=ptr, loc=A.DEVICE)
                frame0 += B
                net.forward(None, A.DEVICE)
                out.append(net.train_pair(n_l, A.MODE_CUDA_REF_SYM if sym else A.MODE_CUDA_REF))
    net.close()
    return out


def driver_exe():
    """The driver binary of the Makefile; rebuilt from its single source if the tree came without it."""
    exe = os.path.join(PKG, "aefft_replay")
    if not os.path.exists(exe):
        subprocess.run(["g++", "-O2", "-std=c++11", "-I", os.path.join(ROOT, "include"), "-o", exe,
                        os.path.join(PKG, "tools", "aefft_replay.cpp"), "-L", PKG, "-laefft", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


def test_replay_driver_matches_the_same_calls_through_ctypes(ctx, tmp_path):
    exe = driver_exe()
    (tmp_path / "New_Layer_Param.txt").write_text(PARAM)
    (tmp_path / "weights").mkdir()
    run = subprocess.run([exe, "--frames", str(B), "--size", f"{NX}x{NY}", "--channels", str(D), "--seed", "1234", "--param",
                          str(tmp_path / "New_Layer_Param.txt"), "--weights", str(tmp_path / "weights"), "--script", SCRIPT],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    lines = run.stdout.splitlines()
    assert lines[-1] == "replay ok"
    assert "Added new layer L 2" in lines and "Deleted last layer" in lines and "Symmetric weights 1" in lines
    got = [float(l.split()[1]) for l in lines if l.startswith("mse ")]
    ctx.set_precision(A.PRECISION_BF16X3)
    try:
        want = replica(ctx, tmp_path)
    finally:
        ctx.set_precision(A.PRECISION_FP32)
    assert len(got) == len(want) == 13
    assert all(np.float32(g) == np.float32(w) for g, w in zip(got, want)), (got, want)
    assert all(np.isfinite(got))
    assert len(list((tmp_path / "weights").iterdir())) == 2


def test_replay_driver_reports_errors(tmp_path):
    exe = driver_exe()
    run = subprocess.run([exe, "--size", "32x32", "--param", str(tmp_path / "missing.txt"), "--script", "n"],
                         capture_output=True, text=True, timeout=120)
    assert run.returncode == 1 and "aefft_replay" in run.stderr


# ---------------------------------------------------------------------------------------------- against the oracle
def _run(exe, tmp_path, args, script, timeout=300):
    (tmp_path / "New_Layer_Param.txt").write_text(PARAM)
    (tmp_path / "weights").mkdir(exist_ok=True)
    run = subprocess.run([exe, "--param", str(tmp_path / "New_Layer_Param.txt"), "--weights", str(tmp_path / "weights"), *args,
                          "--script", script], capture_output=True, text=True, timeout=timeout)
    assert run.returncode == 0, run.stderr
    lines = run.stdout.splitlines()
    assert lines[-1] == "replay ok"
    return lines


def _oracle_net(seed, D, n_pairs):
    """The weights the driver draws: srand(seed), per pair Init_conv(c,b) then Init_conv(f,p) then the four zero-inits that
    still consume rand() (autoencoder.cpp:100-107, :412-428), with the parameters of PARAM."""
    dM, Lk, Ll, scal, rmax = [float(l.split()[1]) for l in PARAM.strip().splitlines()]
    dM, Nk, Nl, scal = int(dM), 2 * (int(Lk) + 1) + 1, 2 * (int(Ll) + 1) + 1, int(scal)
    rng = O.GlibcRand(seed)
    encs, d = [], D
    for _ in range(n_pairs):
        c, b = O.init_conv(rng, dM, d, Nk, Nl, rmax)
        f, p = O.init_conv(rng, d, dM, Nk, Nl, rmax)
        for a, bb in ((dM, d), (d, dM), (dM, d), (d, dM)):
            O.init_conv(rng, a, bb, Nk, Nl, 0.0)
        encs.append((c, b, f, p))
        d = dM
    net_c = [e[0] for e in encs] + [e[2] for e in reversed(encs)]
    net_b = [e[1] for e in encs] + [e[3] for e in reversed(encs)]
    return net_c, net_b, [scal] * n_pairs + [-scal] * n_pairs


def test_replay_coordinate_mse_lines_match_the_oracle(tmp_path):
    """`n t3` on one square frame: the driver's mse lines are the oracle's backprop_gpu (all quirks, momentum carried) on
    the oracle's own forward -- not a comparison of the library with itself."""
    from test_coord_gpu import oracle_forward

    lines = _run(driver_exe(), tmp_path, ["--frames", "1", "--size", "32x32", "--channels", "3", "--seed", "77", "--precision", "fp32"],
                 "n t3")
    got = [float(l.split()[1]) for l in lines if l.startswith("mse ")]
    net_c, net_b, scale = _oracle_net(77, 3, 1)
    c, b, f, p = net_c[0].astype(np.float64), net_b[0].astype(np.float64), net_c[1].astype(np.float64), net_b[1].astype(np.float64)
    st = {k: np.zeros_like(v) for k, v in (("dc", c), ("db", b), ("df", f), ("dp", p))}
    want = []
    for it in range(3):
        x = O.synth_frames(1234, 1, 3, 32, 32, b0=it)
        L = oracle_forward(x, [c, f], [b, p], scale)
        z = lambda a: np.zeros_like(a)
        r = O.backprop_gpu(L[1][0], L[3][0], L[2][0], c, b, f, p, st["dc"], st["db"], st["df"], st["dp"], z(c), z(b), z(f), z(p),
                           0.2, 0.9, quirks=True)
        c, b, f, p = r["c"], r["b"], r["f"], r["p"]
        st = {k: r[k] for k in st}
        want.append(r["mse"])
    assert len(got) == 3 and np.allclose(got, want, rtol=1e-4), (got, want)


def test_replay_fft_mode_matches_the_oracle(tmp_path):
    """`f g t2`: momentum-space forward + backprop_fft of the active pair; the "mse fft:" / "n: .. mse:" lines are the
    oracle's (autoenc_fft + backprop_fft, fft_backproplib.cu:1331-1511) for the weights the driver drew."""
    iters = 6
    lines = _run(driver_exe(), tmp_path, ["--frames", "2", "--size", "32x32", "--channels", "3", "--seed", "91", "--fft-iters", str(iters),
                                          "--del", "0.05"], "n f g t2")
    assert "fft 1" in lines and "fft_l 1" in lines
    got0 = [float(l.split()[2]) for l in lines if l.startswith("mse fft:")]
    gotn = [float(l.split()[3]) for l in lines if l.startswith("n: ")]
    assert len(got0) == 2 and len(gotn) == 2 * iters
    net_c, net_b, scale = _oracle_net(91, 3, 1)
    N, n_l = 2, 0
    want0, wantn = [], []
    for it in range(2):
        x = O.synth_frames(1234, 2, 3, 32, 32, b0=2 * it)
        layers = [O.autoenc_fft(x[k], net_c, net_b, scale, None, 1)[0] for k in range(2)]
        li, lo = 2 * n_l + 1, 2 * N - 1 - 2 * n_l
        inp = np.stack([layers[k][li] for k in range(2)])
        out = np.stack([layers[k][lo] for k in range(2)])
        r = O.backprop_fft(inp, inp, out, net_c[n_l], net_c[N - 1 - n_l], net_b[n_l], net_b[N - 1 - n_l], 0.05, 0, iters)
        net_c[n_l], net_c[N - 1 - n_l], net_b[n_l], net_b[N - 1 - n_l] = r["c"], r["f"], r["b"], r["p"]
        want0.append(r["mse"][0])
        wantn += list(r["mse"][1:])
    assert np.allclose(got0, want0, rtol=2e-4), (got0, want0)
    assert np.allclose(gotn, wantn, rtol=2e-4), (gotn, wantn)


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def test_replay_two_ranks_equal_one_rank_with_twice_the_frames(tmp_path):
    """C++ only, no Python in the training processes: two aefft_replay ranks (one GPU each, the engine's own NCCL
    communicator bootstrapped through a file) print the mse lines of one rank that holds both ranks' frames."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    exe = driver_exe()
    (tmp_path / "New_Layer_Param.txt").write_text(PARAM)
    (tmp_path / "weights").mkdir()
    common = ["--size", "64x64", "--channels", "3", "--seed", "5", "--param", str(tmp_path / "New_Layer_Param.txt"), "--weights",
              str(tmp_path / "weights")]
    script = "n p t3 z t2 f t1"
    one = subprocess.run([exe, "--frames", "4", *common, "--fft-iters", "4", "--script", script], capture_output=True, text=True,
                         timeout=300)
    assert one.returncode == 0, one.stderr
    procs = [subprocess.Popen([exe, "--frames", "2", *common, "--fft-iters", "4", "--device", str(r), "--rank", str(r), "--world", "2",
                               "--id-file", str(tmp_path / "nccl.id"), "--script", script], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se
    pick = lambda text: [float(l.split()[-1]) for l in text.splitlines() if l.startswith(("mse ", "mse fft:", "n: "))]
    want, r0, r1 = pick(one.stdout), pick(outs[0][0]), pick(outs[1][0])
    assert len(want) == 5 + 5 and r0 == r1          # replicas stay identical without any broadcast
    assert np.allclose(r0, want, rtol=2e-5), (r0, want)


def test_momentum_sidecar_makes_a_reload_bit_identical(ctx, tmp_path):
    """Train 2 steps, save weights + the momentum sidecar, train 2 more steps; a second net that loads both and trains the
    same 2 steps ends bit-identical -- with the weight files alone (what the reference saves) the momentum restarts from
    zero and the result differs."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import weights as W

    def fresh():
        ctypes.CDLL("libc.so.6").srand(3)
        net = A.Net(ctx, 3, 32, 32, 2)
        net.add_layer(6, 1, 1, 2, 0.3)
        return net

    frames = [O.synth_frames(8, 2, 3, 32, 32, b0=2 * k) for k in range(4)]
    (tmp_path / "w").mkdir()
    a = fresh()
    for k in range(2):
        a.step(frames[k], A.MODE_CUDA_REF)
    for n, io in ((0, 0), (1, 1)):
        c, b = a.get_conv(n)
        A.saveload_conv(tmp_path / "w", c, b, a.conv_dims(n)[4], 0, io, True)
    a.saveload_momentum(tmp_path / "w", 0, True)
    for k in range(2, 4):
        a.step(frames[k], A.MODE_CUDA_REF)
    want = [a.get_conv(n) for n in range(2)]
    a.close()
    mom, meta = W.read_momentum(next((tmp_path / "w").glob("*.mom")))
    assert meta["dM"] == 6 and mom["dc"].shape == (6, 3, 5, 5) and np.abs(mom["dc"]).max() > 0

    def reload(with_sidecar):
        net = fresh()
        for n, io in ((0, 0), (1, 1)):
            dM, dD, Nk, Nl, sc = net.conv_dims(n)
            c, b = np.empty((dM, dD, Nk, Nl), np.float32), np.empty(dM, np.float32)
            A.saveload_conv(tmp_path / "w", c, b, sc, 0, io, False)
            net.set_conv(n, c, b)
        if with_sidecar:
            net.saveload_momentum(tmp_path / "w", 0, False)
        for k in range(2, 4):
            net.step(frames[k], A.MODE_CUDA_REF)
        got = [net.get_conv(n) for n in range(2)]
        net.close()
        return got

    got = reload(True)
    assert all(np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) for g, w in zip(got, want))
    cold = reload(False)
    assert not np.array_equal(cold[0][0], want[0][0])


def test_replay_raw_video_front_end(ctx, tmp_path):
    """--video / --dump: frames from a raw 8-bit video file (interleaved [frame][Ny][Nx][3], the layout of the
    reference's webcam frames) through ImageToSpin_C on the device, the reconstruction back through SpinToImage_C -- against
    the same calls through ctypes on the same bytes."""
    exe = driver_exe()
    (tmp_path / "New_Layer_Param.txt").write_text(PARAM)
    (tmp_path / "weights").mkdir()
    Bv, nx, ny = 2, 48, 32
    rng = np.random.default_rng(3)
    video = rng.integers(0, 256, size=(5, ny, nx, 3), dtype=np.uint8)  # 5 frames: the second batch wraps around
    (tmp_path / "in.raw").write_bytes(video.tobytes())
    run = subprocess.run([exe, "--frames", str(Bv), "--size", f"{nx}x{ny}", "--seed", "77", "--param",
                          str(tmp_path / "New_Layer_Param.txt"), "--weights", str(tmp_path / "weights"), "--video",
                          str(tmp_path / "in.raw"), "--dump", str(tmp_path / "out.raw"), "--script", "n t3"],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    got_mse = [float(l.split()[1]) for l in run.stdout.splitlines() if l.startswith("mse ")]
    got_img = np.frombuffer((tmp_path / "out.raw").read_bytes(), np.uint8).reshape(Bv, ny, nx, 3)
    # the same through ctypes
    ctypes.CDLL("libc.so.6").srand(77)
    ctx.set_precision(A.PRECISION_BF16X3)
    dM, Lk, Ll, scal, rmax = A.load_param(tmp_path / "New_Layer_Param.txt")
    net = A.Net(ctx, 3, nx, ny, Bv)
    try:
        net.add_layer(dM, Lk, Ll, scal, rmax)
        want_mse = []
        for it in range(3):
            idx = [(it * Bv + b) % len(video) for b in range(Bv)]
            net.set_frames_u8(np.ascontiguousarray(video[idx]))
            net.forward(None, loc=A.DEVICE)
            want_mse.append(net.train_pair(0, A.MODE_CUDA_REF, quirks=0))
        net.forward(None, loc=A.DEVICE)
        ctx.sync()
        want_img = net.get_layer_u8(net.num_layers - 1, 0)
    finally:
        net.close()
    assert np.allclose(got_mse, want_mse, rtol=1e-6), (got_mse, want_mse)
    assert np.array_equal(got_img, want_img)
