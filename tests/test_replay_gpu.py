"""The headless replay driver (autoencoder-fft_b200/tools/aefft_replay.cpp, built by the Makefile) walks the reference
application's key-driven state machine -- add / delete layer from New_Layer_Param.txt, layer cycling with momentum
restart, weight re-draw, symmetric toggle, save / load -- over the C ABI.  The same event sequence is replayed here
through ctypes on the same library; seeded weights, synthetic frames and kernels are deterministic, so every "mse" line
must agree to the printed precision."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import aefft_ctypes as A
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
