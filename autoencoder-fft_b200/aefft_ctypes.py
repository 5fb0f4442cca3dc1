"""ctypes binding of libaefft.so (include/aefft.h) used by tests/, bench.py and __graft_entry__.py.

This is plumbing, not a compute path: every function below forwards to the C ABI, which fails with
AEFFT_ERR_CUDA when no B200 is usable -- there is no CPU fallback anywhere in this package.
Host data are C-contiguous float32 numpy arrays (loc=HOST) or torch CUDA tensors (loc=DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libaefft.so")

OK, ERR_ARG, ERR_CUDA, ERR_IO, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
HOST, DEVICE = 0, 1
CONV_CUDA, CONV_CPU = 0, 1
MODE_CPU_REF, MODE_CUDA_REF, MODE_CUDA_REF_SYM = 0, 1, 2
QUIRK_C1, QUIRK_C3, QUIRK_C4, QUIRKS_ALL = 1, 2, 4, 7
PRECISION_FP32, PRECISION_BF16X3, PRECISION_BF16 = 0, 1, 2

FP = C.POINTER(C.c_float)
_lib = None


class AefftError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"aefft error {code}: {msg}")
        self.code = code


def lib():
    """Load libaefft.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.aefft_last_error.restype = C.c_char_p
        _lib.aefft_launch_count.restype = C.c_int64
        _lib.aefft_launch_count.argtypes = [C.c_void_p]
        _lib.aefft_coord_gbuf_len.restype = C.c_int64
        _lib.aefft_stream.restype = C.c_void_p
        _lib.aefft_stream.argtypes = [C.c_void_p]
    return _lib


def _chk(code):
    if code != OK:
        raise AefftError(code, lib().aefft_last_error().decode())


def _ptr(a):
    """float* of a numpy array / torch tensor / None."""
    if a is None:
        return C.cast(None, FP)
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float32 and a.flags.c_contiguous, "need C-contiguous float32"
        return a.ctypes.data_as(FP)
    if isinstance(a, DevBuf):
        return C.cast(a.ptr, FP)
    if isinstance(a, int):
        return C.cast(a, FP)
    # torch tensor (device or pinned host)
    assert a.is_contiguous() and str(a.dtype) == "torch.float32"
    return C.cast(a.data_ptr(), FP)


def f32(a):
    return np.ascontiguousarray(np.asarray(a, np.float32))


class DevBuf:
    """A float32 device buffer owned through aefft_malloc (tests/bench keep data resident without torch)."""

    def __init__(self, ctx, shape):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.nbytes = int(np.prod(self.shape)) * 4
        self.ptr = ctx.malloc(self.nbytes)

    def data_ptr(self):
        return self.ptr

    @property
    def ndim(self):
        return len(self.shape)

    def numpy(self):
        out = np.empty(self.shape, np.float32)
        self.ctx.memcpy(out.ctypes.data, self.ptr, self.nbytes, 1)
        return out

    def free(self):
        if self.ptr:
            self.ctx.free(self.ptr)
            self.ptr = 0


class Ctx:
    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        _chk(lib().aefft_create(C.byref(self.h), int(device)))

    def close(self):
        if self.h:
            lib().aefft_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _chk(lib().aefft_sync(self.h))

    def set_precision(self, precision: int):
        _chk(lib().aefft_set_precision(self.h, int(precision)))

    @property
    def precision(self) -> int:
        return int(lib().aefft_get_precision(self.h))

    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _chk(lib().aefft_malloc(self.h, C.byref(p), C.c_int64(nbytes)))
        return int(p.value)

    def free(self, ptr: int):
        _chk(lib().aefft_free(self.h, C.c_void_p(ptr)))

    def memcpy(self, dst: int, src: int, nbytes: int, kind: int):
        """kind 0 H2D, 1 D2H, 2 D2D; dst/src are raw addresses."""
        _chk(lib().aefft_memcpy(self.h, C.c_void_p(dst), C.c_void_p(src), C.c_int64(nbytes), kind))

    def to_device(self, a: np.ndarray) -> "DevBuf":
        a = f32(a)
        buf = DevBuf(self, a.shape)
        self.memcpy(buf.ptr, a.ctypes.data, a.nbytes, 0)
        return buf

    def set_gradient_hook(self, fn):
        """fn(dev_ptr: int, n_floats: int) -> None averages the raw momentum-space gradient block over the ranks in place
        (work ordered on the ctx stream); None removes the hook (aefft_set_gradient_hook)."""
        HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)
        if fn is None:
            self._hook = None
            _chk(lib().aefft_set_gradient_hook(self.h, C.cast(None, HOOK), None))
            return

        def tramp(user, ptr, n):
            try:
                fn(int(ptr), int(n))
                return 0
            except Exception:  # the C side turns a non-zero return into an AefftError
                import traceback

                traceback.print_exc()
                return 1

        self._hook = HOOK(tramp)  # keep the trampoline alive
        _chk(lib().aefft_set_gradient_hook(self.h, self._hook, None))

    def set_bin_shard(self, rank: int, world: int):
        """Frequency-bin sharding of backprop_fft (aefft_set_bin_shard); the gradient hook must then SUM over devices."""
        _chk(lib().aefft_set_bin_shard(self.h, int(rank), int(world)))

    # ---------------------------------------------------------------- multi-GPU (the engine's own NCCL communicator)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _chk(lib().aefft_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, id_bytes: bytes, rank: int, world: int):
        assert len(id_bytes) == 128
        _chk(lib().aefft_comm_init(self.h, C.c_char_p(id_bytes), int(rank), int(world)))

    def comm_destroy(self):
        _chk(lib().aefft_comm_destroy(self.h))

    def comm_allreduce(self, dev_ptr: int, n: int, op: int = 0):
        _chk(lib().aefft_comm_allreduce(self.h, C.c_void_p(dev_ptr), C.c_int64(n), int(op)))

    def profile_enable(self, on: bool):
        """Bracket every kernel launch of this ctx with a CUDA event pair (aefft_profile_enable)."""
        _chk(lib().aefft_profile_enable(self.h, 1 if on else 0))

    def profile_records(self, max_rows: int = 64):
        """Synchronise, aggregate by kernel name and clear: [{name, ms, launches, flops, bytes}] (algorithmic work)."""
        names = C.create_string_buffer(64 * max_rows)
        kms = (C.c_float * max_rows)()
        cnt = (C.c_int64 * max_rows)()
        fl = (C.c_double * max_rows)()
        by = (C.c_double * max_rows)()
        nrows = C.c_int()
        _chk(lib().aefft_profile_read(self.h, max_rows, names, kms, cnt, fl, by, C.byref(nrows)))
        rows = []
        for k in range(nrows.value):
            nm = names.raw[64 * k: 64 * k + 64].split(b"\0")[0].decode()
            rows.append(dict(name=nm, ms=float(kms[k]), launches=int(cnt[k]), flops=float(fl[k]), bytes=float(by[k])))
        return rows

    @property
    def launches(self) -> int:
        return int(lib().aefft_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(lib().aefft_stream(self.h) or 0)

    # ---------------------------------------------------------------- forward
    def conv_fwd(self, x, c, b, convention=CONV_CUDA, out=None, loc=HOST):
        dM, dD, Nk, Nl = c.shape
        B = x.shape[0] if x.ndim == 4 else 1
        Nx, Ny = x.shape[-2:]
        if out is None:
            out = np.empty(((B,) if x.ndim == 4 else ()) + (dM, Nx, Ny), np.float32)
        _chk(lib().aefft_conv_fwd(self.h, loc, convention, C.c_int64(B), dD, dM, Nx, Ny, Nk, Nl, _ptr(x), _ptr(c),
                                  _ptr(b), _ptr(out)))
        return out

    def pool(self, x, scale, out_shape, out=None, loc=HOST):
        B = x.shape[0] if x.ndim == 4 else 1
        D, Nx, Ny = x.shape[-3:]
        oNx, oNy = out_shape
        if out is None:
            out = np.empty(((B,) if x.ndim == 4 else ()) + (D, oNx, oNy), np.float32)
        _chk(lib().aefft_pool(self.h, loc, C.c_int64(B), D, Nx, Ny, oNx, oNy, scale, _ptr(x), _ptr(out)))
        return out

    def portion(self, x, q, loc=HOST):
        B = x.shape[0] if x.ndim == 4 else 1
        D, Nx, Ny = x.shape[-3:]
        out = np.empty(((B,) if x.ndim == 4 else ()) + (D, Nx // q, Ny // q), np.float32)
        _chk(lib().aefft_portion(self.h, loc, C.c_int64(B), D, Nx, Ny, q, _ptr(x), _ptr(out)))
        return out

    def synth_frames(self, seed, B, D, Nx, Ny, b0=0, out=None, loc=HOST):
        if out is None:
            out = np.empty((B, D, Nx, Ny), np.float32)
        _chk(lib().aefft_synth_frames(self.h, loc, C.c_uint64(seed), C.c_int64(b0), C.c_int64(B), D, Nx, Ny, _ptr(out)))
        return out

    # ---------------------------------------------------------------- coordinate training
    def backprop_coord(self, mode, inp, out, hin, c, b, f, p, dc=None, db=None, df=None, dp=None, ddc=None, ddb=None,
                       ddf=None, ddp=None, delmax=0.2, alpha=0.9, active=1, quirks=QUIRKS_ALL, loc=HOST):
        """In-place on c,b,f,p and the momentum / last-gradient buffers; returns the printed mse."""
        dM, dD, Nk, Nl = c.shape
        B = inp.shape[0] if inp.ndim == 4 else 1
        Nx, Ny = inp.shape[-2:]
        mse = C.c_float(0)
        _chk(lib().aefft_backprop_coord(self.h, loc, mode, quirks, C.c_int64(B), dD, dM, Nx, Ny, Nk, Nl, _ptr(inp),
                                        _ptr(out), _ptr(hin), _ptr(c), _ptr(b), _ptr(f), _ptr(p), _ptr(dc), _ptr(db),
                                        _ptr(df), _ptr(dp), _ptr(ddc), _ptr(ddb), _ptr(ddf), _ptr(ddp),
                                        C.c_float(delmax), C.c_float(alpha), int(active), C.byref(mse)))
        return float(mse.value)

    def coord_gradients(self, mode, quirks, B, dD, dM, Nx, Ny, Nk, Nl, inp, out, hin, c, f, gbuf):
        _chk(lib().aefft_coord_gradients(self.h, mode, quirks, C.c_int64(B), dD, dM, Nx, Ny, Nk, Nl, _ptr(inp),
                                         _ptr(out), _ptr(hin), _ptr(c), _ptr(f), _ptr(gbuf)))

    def coord_update(self, mode, B_global, dD, dM, Nx, Ny, Nk, Nl, gbuf, c, b, f, p, dc, db, df, dp, ddc, ddb, ddf, ddp,
                     delmax, alpha, mse_dev=None):
        _chk(lib().aefft_coord_update(self.h, mode, C.c_int64(B_global), dD, dM, Nx, Ny, Nk, Nl, _ptr(gbuf), _ptr(c),
                                      _ptr(b), _ptr(f), _ptr(p), _ptr(dc), _ptr(db), _ptr(df), _ptr(dp), _ptr(ddc),
                                      _ptr(ddb), _ptr(ddf), _ptr(ddp), C.c_float(delmax), C.c_float(alpha),
                                      _ptr(mse_dev)))

    # ---------------------------------------------------------------- momentum space
    def fft_r2c(self, x, loc=HOST, out=None):
        Nx, Ny = x.shape[-2:]
        batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
        if out is None:
            out = np.empty(tuple(x.shape[:-2]) + (Nx, Ny // 2 + 1, 2), np.float32)
        _chk(lib().aefft_fft_r2c(self.h, loc, C.c_int64(batch), Nx, Ny, _ptr(x), _ptr(out)))
        return out

    def fft_c2r(self, spec, Ny, loc=HOST, out=None):
        Nx = spec.shape[-3]
        batch = int(np.prod(spec.shape[:-3])) if spec.ndim > 3 else 1
        if out is None:
            out = np.empty(tuple(spec.shape[:-3]) + (Nx, Ny), np.float32)
        _chk(lib().aefft_fft_c2r(self.h, loc, C.c_int64(batch), Nx, Ny, _ptr(spec), _ptr(out)))
        return out

    def kernel_pad(self, c, Nx, Ny):
        dM, dD, Nk, Nl = c.shape
        out = np.empty((dM, dD, Nx, Ny), np.float32)
        _chk(lib().aefft_kernel_pad(self.h, HOST, dM, dD, Nk, Nl, Nx, Ny, _ptr(c), _ptr(out)))
        return out

    def kernel_spectrum(self, c, Nx, Ny):
        dM, dD, Nk, Nl = c.shape
        out = np.empty((dM, dD, Nx, Ny // 2 + 1, 2), np.float32)
        _chk(lib().aefft_kernel_spectrum(self.h, HOST, dM, dD, Nk, Nl, Nx, Ny, _ptr(c), _ptr(out)))
        return out

    def autoenc_fft(self, x, net_c, net_b, scale, layer_shapes, cfreq=None, fft_l=1):
        """x [B,D,Nx,Ny] or [D,Nx,Ny]; layer_shapes: list of (D,Nx,Ny) for all 2*n_conv+1 layers.
        Returns (layers list, cfreq list).  cfreq: list of wire-format spectra (valid cache) or None."""
        x = f32(x)
        batched = x.ndim == 4
        B = x.shape[0] if batched else 1
        n_conv = len(net_c)
        dims = np.array([d for c in net_c for d in c.shape], np.int32)
        c_all = np.concatenate([f32(c).ravel() for c in net_c])
        b_all = np.concatenate([f32(b).ravel() for b in net_b])
        coff = np.cumsum([0] + [c.size for c in net_c[:-1]]).astype(np.int64)
        boff = np.cumsum([0] + [b.size for b in net_b[:-1]]).astype(np.int64)
        ldims = np.array([d for s in layer_shapes for d in s], np.int32)
        lsz = [int(np.prod(s)) for s in layer_shapes]
        loff = np.cumsum([0] + lsz[:-1]).astype(np.int64)
        lstride = int(sum(lsz))
        layers_all = np.zeros((B, lstride), np.float32)
        layers_all[:, : lsz[0]] = x.reshape(B, -1)
        # spectra sizes depend on the resolution each conv runs at
        cf_sizes = []
        for n in range(n_conv):
            lay = layer_shapes[2 * n + 1] if n < n_conv // 2 else layer_shapes[2 * n]
            dM, dD = net_c[n].shape[:2]
            cf_sizes.append(dM * dD * lay[1] * (lay[2] // 2 + 1) * 2)
        cfoff = np.cumsum([0] + cf_sizes[:-1]).astype(np.int64)
        cfreq_all = np.zeros(int(sum(cf_sizes)), np.float32)
        valid = 0
        if cfreq is not None:
            valid = 1
            for n in range(n_conv):
                cfreq_all[cfoff[n] : cfoff[n] + cf_sizes[n]] = f32(cfreq[n]).ravel()
        scale_a = np.array(scale, np.int32)
        I32 = C.POINTER(C.c_int32)
        I64 = C.POINTER(C.c_int64)
        _chk(lib().aefft_autoenc_fft(self.h, HOST, C.c_int64(B), n_conv, dims.ctypes.data_as(I32), _ptr(c_all),
                                     coff.ctypes.data_as(I64), _ptr(b_all), boff.ctypes.data_as(I64),
                                     scale_a.ctypes.data_as(I32), len(layer_shapes), ldims.ctypes.data_as(I32),
                                     _ptr(layers_all), loff.ctypes.data_as(I64), C.c_int64(lstride), valid,
                                     _ptr(cfreq_all), cfoff.ctypes.data_as(I64), int(fft_l)))
        layers = []
        for l, s in enumerate(layer_shapes):
            a = layers_all[:, loff[l] : loff[l] + lsz[l]].reshape((B,) + tuple(s))
            layers.append(a if batched else a[0])
        spectra = [cfreq_all[cfoff[n] : cfoff[n] + cf_sizes[n]].copy() for n in range(n_conv)]
        return layers, spectra

    def backprop_fft(self, inp, expout, out, c, f, b, p, del0, maxdiff=0, n_iter=100, cfreq=None, ffreq=None, loc=HOST):
        """In-place on c,f,b,p (and cfreq/ffreq when given); returns the mse trace (n_iter+1 values)."""
        dM, dD, Nk, Nl = c.shape
        B = inp.shape[0] if inp.ndim == 4 else 1
        Nx, Ny = inp.shape[-2:]
        trace = np.zeros(n_iter + 1, np.float32)
        _chk(lib().aefft_backprop_fft(self.h, loc, C.c_int64(B), dD, dM, Nx, Ny, Nk, Nl, _ptr(inp), _ptr(expout),
                                      _ptr(out), _ptr(cfreq), _ptr(c), _ptr(ffreq), _ptr(f), _ptr(b), _ptr(p),
                                      C.c_float(del0), int(maxdiff), int(n_iter), _ptr(trace)))
        return trace


# -------------------------------------------------------------------- glue (host only, no GPU needed)
def init_conv(mS, dD, kS, lS, rmax):
    c = np.zeros((mS, dD, kS, lS), np.float32)
    b = np.zeros((mS,), np.float32)
    _chk(lib().aefft_init_conv(_ptr(c), _ptr(b), mS, dD, kS, lS, C.c_float(rmax)))
    return c, b


def saveload_conv(directory, c, b, scale, L, io, write):
    dM, dD, Nk, Nl = c.shape
    _chk(lib().aefft_saveload_conv(str(directory).encode(), _ptr(c), _ptr(b), dM, dD, Nk, Nl, scale, L, io, write))


def load_param(path):
    dM, Lk, Ll, scal = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rmax = C.c_float()
    _chk(lib().aefft_load_param(str(path).encode(), C.byref(dM), C.byref(Lk), C.byref(Ll), C.byref(scal), C.byref(rmax)))
    return dM.value, Lk.value, Ll.value, scal.value, rmax.value


class Net:
    """Device-resident network (aefft_net_*): the headless replay of autoencoder.cpp's state model."""

    def __init__(self, ctx: Ctx, D, Nx, Ny, B):
        self.ctx = ctx
        self.B = B
        self.h = C.c_void_p()
        _chk(lib().aefft_net_create(ctx.h, C.byref(self.h), D, Nx, Ny, C.c_int64(B)))

    def close(self):
        if self.h:
            lib().aefft_net_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_layer(self, dM, Lk, Ll, scal, rmax):
        _chk(lib().aefft_net_add_layer(self.h, dM, Lk, Ll, scal, C.c_float(rmax)))

    def delete_layer(self):
        _chk(lib().aefft_net_delete_layer(self.h))

    @property
    def num_pairs(self):
        return lib().aefft_net_num_pairs(self.h)

    @property
    def num_layers(self):
        return lib().aefft_net_num_layers(self.h)

    def conv_dims(self, n):
        v = [C.c_int() for _ in range(5)]
        _chk(lib().aefft_net_conv_dims(self.h, n, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)  # dM, dD, Nk, Nl, scale

    def get_conv(self, n):
        dM, dD, Nk, Nl, _ = self.conv_dims(n)
        c = np.empty((dM, dD, Nk, Nl), np.float32)
        b = np.empty((dM,), np.float32)
        _chk(lib().aefft_net_get_conv(self.h, n, _ptr(c), _ptr(b)))
        return c, b

    def set_conv(self, n, c, b):
        _chk(lib().aefft_net_set_conv(self.h, n, _ptr(f32(c)), _ptr(f32(b))))

    def set_symmetric(self, n_l):
        _chk(lib().aefft_net_set_symmetric(self.h, n_l))

    def reset_momentum(self, n_l):
        _chk(lib().aefft_net_reset_momentum(self.h, n_l))

    def layer_info(self, l):
        D, Nx, Ny = C.c_int(), C.c_int(), C.c_int()
        p = FP()
        _chk(lib().aefft_net_layer(self.h, l, C.byref(D), C.byref(Nx), C.byref(Ny), C.byref(p)))
        return D.value, Nx.value, Ny.value, C.cast(p, C.c_void_p).value

    def layer(self, l):
        """Copy layer l to the host as [B,D,Nx,Ny]."""
        D, Nx, Ny, ptr = self.layer_info(l)
        out = np.empty((self.B, D, Nx, Ny), np.float32)
        self.ctx.memcpy(out.ctypes.data, ptr, out.nbytes, 1)
        return out

    def forward(self, frames, loc=HOST):
        _chk(lib().aefft_net_forward(self.h, loc, _ptr(frames)))

    def train_pair(self, n_l, mode, delmax=0.2, alpha=0.9, quirks=QUIRKS_ALL, want_mse=True):
        mse = C.c_float(0)
        _chk(lib().aefft_net_train_pair(self.h, n_l, mode, quirks, C.c_float(delmax), C.c_float(alpha),
                                        C.byref(mse) if want_mse else None))
        return float(mse.value)

    def pair_gradients(self, n_l, mode, quirks=QUIRKS_ALL):
        p = FP()
        n = C.c_int64()
        _chk(lib().aefft_net_pair_gradients(self.h, n_l, mode, quirks, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, n.value

    def pair_update(self, n_l, mode, B_global, delmax=0.2, alpha=0.9, want_mse=False):
        mse = C.c_float(0)
        _chk(lib().aefft_net_pair_update(self.h, n_l, mode, C.c_int64(B_global), C.c_float(delmax), C.c_float(alpha),
                                         C.byref(mse) if want_mse else None))
        return float(mse.value)

    # ---------------------------------------------------------------- momentum space on the resident net
    def fft_forward(self, frames, fft_l=1, loc=HOST):
        _chk(lib().aefft_net_fft_forward(self.h, loc, _ptr(frames), int(fft_l)))

    def fft_step(self, frames, del0=0.2, maxdiff=0, n_iter=1, fft_l=0, loc=HOST, want_mse=True):
        """Returns the mse traces [pairs][n_iter+1] (or None)."""
        trace = np.zeros((self.num_pairs, n_iter + 1), np.float32) if want_mse else None
        _chk(lib().aefft_net_fft_step(self.h, loc, _ptr(frames), C.c_float(del0), int(maxdiff), int(n_iter), int(fft_l),
                                      _ptr(trace)))
        return trace

    def fft_train_pair(self, n_l, del0=0.2, maxdiff=0, n_iter=100):
        trace = np.zeros(n_iter + 1, np.float32)
        _chk(lib().aefft_net_fft_train_pair(self.h, int(n_l), C.c_float(del0), int(maxdiff), int(n_iter), _ptr(trace)))
        return trace

    def get_cfreq(self, n):
        dM, dD, Nk, Nl, _ = self.conv_dims(n)
        N = 2 * self.num_pairs
        _, Nx, Ny, _ = self.layer_info(2 * n + 1 if n < N // 2 else 2 * n)
        out = np.empty((dM, dD, Nx, Ny // 2 + 1, 2), np.float32)
        _chk(lib().aefft_net_get_cfreq(self.h, n, _ptr(out), C.c_int64(out.size)))
        return out

    def saveload_momentum(self, directory, n_l, write):
        _chk(lib().aefft_net_saveload_momentum(self.h, str(directory).encode(), int(n_l), 1 if write else 0))

    def fused_layout(self, mode):
        """(offsets per pair, total floats) of the fused raw gradient block a data-parallel step all-reduces once."""
        P = self.num_pairs
        off = (C.c_int64 * P)()
        tot = C.c_int64()
        _chk(lib().aefft_net_fused_layout(self.h, mode, off, C.byref(tot)))
        return [int(v) for v in off], int(tot.value)

    def set_frames_u8(self, images, loc=HOST):
        """ImageToSpin_C on the device: images uint8 [B][Ny][Nx][D] (numpy array, or a raw address with loc)."""
        ptr = images.ctypes.data if isinstance(images, np.ndarray) else int(images)
        _chk(lib().aefft_net_set_frames_u8(self.h, loc, C.c_void_p(ptr)))

    def get_layer_u8(self, layer, mode=0):
        """SpinToImage_C (mode 0) / SpinToImage_V (mode 1) of a layer with <= 4 channels: uint8 [B][Ny][Nx][D]."""
        D, Nx, Ny = self.layer_info(layer)[:3]
        out = np.empty((self.B, Ny, Nx, D), np.uint8)
        _chk(lib().aefft_net_get_layer_u8(self.h, int(layer), int(mode), HOST, C.c_void_p(out.ctypes.data)))
        return out

    def step(self, frames, mode, delmax=0.2, alpha=0.9, quirks=QUIRKS_ALL, loc=HOST, mse=None):
        _chk(lib().aefft_net_step(self.h, loc, _ptr(frames), mode, quirks, C.c_float(delmax), C.c_float(alpha),
                                  _ptr(mse)))
