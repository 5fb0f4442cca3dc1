"""ctypes door onto oracle/_ref/libref.so (the UNMODIFIED reference, built by oracle/Makefile).
TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden*.py and bench.py's reference arm."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
_lib = None
FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)
LP = C.POINTER(C.c_long)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(FP)


def f32(a):
    return np.ascontiguousarray(np.asarray(a, np.float32))


def srand(seed):
    lib().ref_srand(C.c_uint(seed))


def init_conv(mS, dD, kS, lS, rmax):
    c = np.zeros((mS, dD, kS, lS), np.float32)
    b = np.zeros((mS,), np.float32)
    lib().ref_init_conv(_f(c), _f(b), mS, dD, kS, lS, C.c_float(rmax))
    return c, b


def pool(x, scale, out_shape):
    x = f32(x)
    D, Nx, Ny = x.shape
    out = np.zeros((D,) + tuple(out_shape), np.float32)
    lib().ref_pool(_f(x), _f(out), D, Nx, Ny, out_shape[0], out_shape[1], scale)
    return out


def conv_cpu(x, c, b):
    x, c, b = f32(x), f32(c), f32(b)
    dM, dD, Nk, Nl = c.shape
    _, Nx, Ny = x.shape
    out = np.zeros((dM, Nx, Ny), np.float32)
    lib().ref_conv_cpu(_f(x), _f(out), _f(c), _f(b), dD, dM, Nx, Ny, Nk, Nl)
    return out


def conv_gpu(x, c, b):
    x, c, b = f32(x), f32(c), f32(b)
    dM, dD, Nk, Nl = c.shape
    _, Nx, Ny = x.shape
    out = np.zeros((dM, Nx, Ny), np.float32)
    lib().ref_conv_gpu(_f(x), _f(out), _f(c), _f(b), dD, dM, Nx, Ny, Nk, Nl)
    return out


def backprop_cpu(inp, out, hin, c, b, f, p, delta):
    inp, out, hin = f32(inp), f32(out), f32(hin)
    c, b, f, p = (f32(t).copy() for t in (c, b, f, p))
    dM, dD, Nk, Nl = c.shape
    _, Nx, Ny = inp.shape
    lib().ref_backprop_cpu(_f(inp), _f(out), _f(hin), _f(c), _f(b), _f(f), _f(p), C.c_float(delta), dD, dM, Nx,
                           Ny, Nk, Nl)
    return dict(c=c, b=b, f=f, p=p)


def portion(inp, hin, out, q):
    inp, hin, out = f32(inp), f32(hin), f32(out)
    D, Nx, Ny = inp.shape
    M = hin.shape[0]
    a = np.zeros((D, Nx // q, Ny // q), np.float32)
    h = np.zeros((M, Nx // q, Ny // q), np.float32)
    o = np.zeros((D, Nx // q, Ny // q), np.float32)
    lib().ref_portion(_f(inp), _f(hin), _f(out), _f(a), _f(h), _f(o), D, M, Nx, Ny, q)
    return a, h, o


def saveload_conv(c, b, scale, L, io, write):
    c, b = f32(c).copy(), f32(b).copy()
    dM, dD, Nk, Nl = c.shape
    lib().ref_saveload_conv(_f(c), _f(b), dM, dD, Nk, Nl, scale, L, io, write)
    return c, b


def loadparam():
    dM, Lk, Ll, sc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rmax = C.c_float()
    lib().ref_loadparam(C.byref(dM), C.byref(Lk), C.byref(Ll), C.byref(sc), C.byref(rmax))
    return dM.value, Lk.value, Ll.value, sc.value, rmax.value


def backprop_gpu(sym, inp, out, hin, c, b, f, p, dc, db, df, dp, ddc, ddb, ddf, ddp, delmax, alpha, active=1):
    inp, out, hin = f32(inp), f32(out), f32(hin)
    names = "c b f p dc db df dp ddc ddb ddf ddp".split()
    arrs = [f32(t).copy() for t in (c, b, f, p, dc, db, df, dp, ddc, ddb, ddf, ddp)]
    dM, dD, Nk, Nl = arrs[0].shape
    _, Nx, Ny = inp.shape
    lib().ref_backprop_gpu(int(sym), _f(inp), _f(out), _f(hin), *[_f(a) for a in arrs], C.c_float(delmax),
                           C.c_float(alpha), int(active), dD, dM, Nx, Ny, Nk, Nl)
    return dict(zip(names, arrs))


def kernel_pad(c, Nx, Ny):
    c = f32(c)
    dM, dD, Nk, Nl = c.shape
    out = np.zeros((dM, dD, Nx, Ny), np.float32)
    lib().ref_kernel_pad(_f(c), _f(out), dM, dD, Nk, Nl, Nx, Ny)
    return out


def autoenc_fft(x, net_c, net_b, scale, layer_shapes, net_cfreq=None, fft_l=1):
    """layer_shapes: list of (D,Nx,Ny) for the 2*n_conv+1 layers.  Returns (layers list, cfreq list)."""
    n_conv = len(net_c)
    dims = np.array([d for c in net_c for d in c.shape], np.int32)
    c_all = np.concatenate([f32(c).ravel() for c in net_c])
    b_all = np.concatenate([f32(b).ravel() for b in net_b])
    coff = np.cumsum([0] + [c.size for c in net_c])[:-1].astype(np.int64)
    boff = np.cumsum([0] + [len(b) for b in net_b])[:-1].astype(np.int64)
    sc = np.array(scale, np.int32)
    ldims = np.array([d for s in layer_shapes for d in s], np.int32)
    lsz = [int(np.prod(s)) for s in layer_shapes]
    loff = np.cumsum([0] + lsz)[:-1].astype(np.int64)
    layers_all = np.zeros(sum(lsz), np.float32)
    layers_all[: lsz[0]] = f32(x).ravel()
    # spectra sizes: conv n runs at the resolution of its input layer (enc: after pool; dec: before unpool)
    cflen = []
    for n in range(n_conv):
        D, Nx, Ny = layer_shapes[2 * n + 1] if n < n_conv // 2 else layer_shapes[2 * n]
        dM, dD = net_c[n].shape[:2]
        cflen.append(dM * dD * Nx * (Ny // 2 + 1) * 2)
    cfoff = np.cumsum([0] + cflen)[:-1].astype(np.int64)
    cflen = np.array(cflen, np.int64)
    cf_all = np.zeros(int(cflen.sum()), np.float32)
    state = 0
    if net_cfreq is not None:
        state = 1
        for n in range(n_conv):
            cf_all[cfoff[n] : cfoff[n] + cflen[n]] = f32(net_cfreq[n]).ravel()
    lib().ref_autoenc_fft(n_conv, dims.ctypes.data_as(IP), _f(c_all), coff.ctypes.data_as(LP), _f(b_all),
                          boff.ctypes.data_as(LP), sc.ctypes.data_as(IP), len(layer_shapes),
                          ldims.ctypes.data_as(IP), _f(layers_all), loff.ctypes.data_as(LP), state, _f(cf_all),
                          cfoff.ctypes.data_as(LP), cflen.ctypes.data_as(LP), int(fft_l))
    layers = [layers_all[loff[i] : loff[i] + lsz[i]].reshape(layer_shapes[i]).copy() for i in range(len(lsz))]
    cfs = [cf_all[cfoff[n] : cfoff[n] + cflen[n]].copy() for n in range(n_conv)]
    return layers, cfs


def backprop_fft(inp, expout, out, cfreq, c, ffreq, f, b, p, del0, maxdiff=0):
    inp, expout, out = f32(inp), f32(expout), f32(out)
    cfreq, c, ffreq, f, b, p = (f32(t).copy() for t in (cfreq, c, ffreq, f, b, p))
    dM, dD, Nk, Nl = c.shape
    _, Nx, Ny = inp.shape
    lib().ref_backprop_fft(_f(inp), _f(expout), _f(out), _f(cfreq), _f(c), _f(ffreq), _f(f), _f(b), _f(p), dD, dM,
                           Nx, Ny, Nk, Nl, C.c_float(del0), int(maxdiff))
    return dict(cfreq=cfreq, c=c, ffreq=ffreq, f=f, b=b, p=p)
