"""Drop-in check: a C++ driver written against the reference's header names (netlib.h / backproplib.h / fft_backproplib.h,
nested std::vector signatures) is compiled against autoencoder-fft_b200/shim and run on the GPU; its dumps are compared
with the numpy oracle replaying the same call sequence."""
import os
import subprocess

import numpy as np
import pytest

import oracle_np as O
from conftest import ROOT

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "autoencoder-fft_b200")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    exe = d / "shim_driver"
    subprocess.run(["g++", "-O1", "-std=c++11", "-I", os.path.join(PKG, "shim"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "shim", "shim_driver.cpp"), "-o", str(exe), "-L", PKG, "-laefft_shim", "-laefft",
                    f"-Wl,-rpath,{PKG}"], check=True)
    (d / "New_Layer_Param.txt").write_text("Layer_depth 4\nKernel_L_x 1\nKernel_L_y 1\nPooling_scale 2\nMax_Rand_Init 0.3\n")
    (d / "weights").mkdir()
    return d, exe


def take(buf, pos, shape):
    n = int(np.prod(shape))
    return buf[pos:pos + n].reshape(shape), pos + n


def inputs(buf):
    pos = 0
    x, pos = take(buf, pos, (3, 32, 32))
    c, pos = take(buf, pos, (4, 3, 5, 5))
    b, pos = take(buf, pos, (4,))
    f, pos = take(buf, pos, (3, 4, 5, 5))
    p, pos = take(buf, pos, (3,))
    rng = O.GlibcRand(1234)
    oc, ob = O.init_conv(rng, 4, 3, 5, 5, 0.3)
    of, op = O.init_conv(rng, 3, 4, 5, 5, 0.3)
    assert np.array_equal(c, oc) and np.array_equal(b, ob) and np.array_equal(f, of) and np.array_equal(p, op)
    return x, c, b, f, p, pos


def test_shim_coordinate_sequence(driver):
    d, exe = driver
    subprocess.run([str(exe), "coord", str(d / "coord.bin")], check=True, cwd=d, stdout=subprocess.DEVNULL)
    buf = np.fromfile(d / "coord.bin", np.float32)
    x, c, b, f, p, pos = inputs(buf)
    pin = O.pool(x, 2)
    hC = O.conv_gpu(pin, c, b).astype(np.float32)
    phC = O.conv_gpu(hC, f, p).astype(np.float32)
    outl = O.pool(phC, -2, (32, 32))
    for want, shape in ((pin, (3, 16, 16)), (hC, (4, 16, 16)), (phC, (3, 16, 16)), (outl, (3, 32, 32))):
        got, pos = take(buf, pos, shape)
        assert O.rel_l2(got, want) < 5e-5
    z = lambda a: np.zeros_like(a)
    want = O.backprop_gpu_cc(pin, phC, hC, c, b, f, p, z(c), z(b), z(f), z(p), z(c), z(b), z(f), z(p), 0.2, 0.9)
    for key, shape in (("c", c.shape), ("b", b.shape), ("f", f.shape), ("p", p.shape)):
        got, pos = take(buf, pos, shape)
        assert O.rel_l2(got, want[key]) < 1e-4, key
    assert pos == buf.size
    assert os.path.exists(d / "weights" / "C_weights_0_in_D=3_M=4_Lk=1_Ll=1_S=2.conv")  # SaveLoad_conv's file name


def test_shim_fft_sequence(driver):
    d, exe = driver
    subprocess.run([str(exe), "fft", str(d / "fft.bin")], check=True, cwd=d, stdout=subprocess.DEVNULL)
    buf = np.fromfile(d / "fft.bin", np.float32)
    x, c, b, f, p, pos = inputs(buf)
    layers, spectra = O.autoenc_fft(x, [c, f], [b, p], [2, -2], None, 1)
    got_layers = []
    for L in layers:
        got, pos = take(buf, pos, L.shape)
        got_layers.append(got)
        assert O.rel_l2(got, L) < 5e-5
    want = O.backprop_fft(got_layers[1], got_layers[1], got_layers[3], c, f, b, p, 0.005, 0, 100, cfreq=spectra[0], ffreq=spectra[1])
    for key, shape in (("c", c.shape), ("b", b.shape), ("f", f.shape), ("p", p.shape)):
        got, pos = take(buf, pos, shape)
        assert O.rel_l2(got, want[key]) < 1e-4, key
    assert pos == buf.size
