"""CPU oracle for the autoencoder hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement (float64 arithmetic unless stated) of the reference algorithm, function by function,
each citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product path (autoencoder-fft_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against the reference's own compiled code:
  * coordinate / CPU functions (Init_conv, Pool, Conv, backprop, Portion, kernel_pad, SaveLoad_conv):
    against oracle/_ref/libref.so run in the build container -> tests/golden/*.npz (tests/golden/make_golden.py)
    and live in tests/test_oracle_vs_ref.py whenever libref.so is present;
  * CUDA functions (Conv_gpu, backprop_gpu, backprop_gpu_cc, autoenc_fft, backprop_fft): against libref.so
    run on a B200 (tests/golden/make_golden_gpu.py, executed under gpurun; outputs committed).
  * the DFT itself lives in closed-source cuFFT; here it is numpy.fft (pocketfft), mathematically defined.

Layout everywhere: feature maps [ch][Nx][Ny] (j fastest); kernels c[dM][dD][Nk][Nl], f[dD][dM][Nk][Nl]
(SURVEY App. A.1).  Batched inputs carry a leading frame axis [B].
"""
from __future__ import annotations

import numpy as np

F64 = np.float64
F32 = np.float32

# --------------------------------------------------------------------------------------------------
# glibc rand() (TYPE_3 additive feedback) -- what `srand(seed)` + Init_conv (netlib.cpp:167-197) consume.
# --------------------------------------------------------------------------------------------------


class GlibcRand:
    """Bit-exact glibc srand()/rand() (random_r TYPE_3: x[i] = x[i-3] + x[i-31], output >> 1)."""

    RAND_MAX = 2147483647

    def __init__(self, seed: int = 1):
        self.srand(seed)

    def srand(self, seed: int) -> None:
        seed = seed & 0xFFFFFFFF
        if seed == 0:
            seed = 1
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            # 16807 * r[i-1] % 2147483647 computed the glibc way (signed hi/lo split)
            prev = r[i - 1]
            if prev >= 2**31:
                prev -= 2**32  # int32_t word
            hi = int(prev / 127773)  # C division truncates toward zero
            lo = prev - hi * 127773
            word = 16807 * lo - 2836 * hi
            if word < 0:
                word += 2147483647
            r[i] = word & 0xFFFFFFFF
        for i in range(31, 34):
            r[i] = r[i - 31]
        self._r = r
        for _ in range(34, 344):
            self._step()

    def _step(self) -> int:
        r = self._r
        v = (r[-31] + r[-3]) & 0xFFFFFFFF
        r.append(v)
        if len(r) > 64:
            del r[:-34]
        return v

    def rand(self) -> int:
        return self._step() >> 1


def init_conv(rng: GlibcRand, mS: int, dD: int, kS: int, lS: int, rmax: float):
    """netlib.cpp:167-197 Init_conv: draw order m,d,k,l then b[m]; r = -max + 2*max*(float)rand()/(float)RAND_MAX
    evaluated in float32 exactly as the C expression (all operands float)."""
    c = np.zeros((mS, dD, kS, lS), F32)
    b = np.zeros((mS,), F32)
    mx = F32(rmax)
    two_mx = F32(F32(2) * mx)
    rm = F32(GlibcRand.RAND_MAX)  # 2147483648.0f
    for m in range(mS):
        for d in range(dD):
            for k in range(kS):
                for l in range(lS):
                    c[m, d, k, l] = F32(-mx + F32(F32(two_mx * F32(rng.rand())) / rm))
        b[m] = F32(-mx + F32(F32(two_mx * F32(rng.rand())) / rm))
    return c, b


# --------------------------------------------------------------------------------------------------
# tap offsets and shifted views
# --------------------------------------------------------------------------------------------------


def tap_offsets(N: int, path: str) -> np.ndarray:
    """Offsets ik(k): tap k multiplies in[i - ik(k)].
    'cpu'  netlib.cpp:325,338 / 371,391:  ak=(Nk-1)/2-1,      ik=-2ak-1+k
    'cuda' backproplib.cu:123,89 / 301,369: ak=((Nk-1)/2-1)/2, ik=-2ak-1+k
    'fft'  fft_backproplib.cu:1034-1057, 579-598: ik=k-Nk/2 (circular)."""
    k = np.arange(N)
    if path == "cpu":
        ak = (N - 1) // 2 - 1
        return -2 * ak - 1 + k
    if path == "cuda":
        ak = ((N - 1) // 2 - 1) // 2
        return -2 * ak - 1 + k
    if path == "fft":
        return k - N // 2
    raise ValueError(path)


def shifted(x: np.ndarray, si: int, sj: int, lo: int = 0) -> np.ndarray:
    """y[..., i, j] = x[..., i-si, j-sj] where lo <= i-si < Nx and lo <= j-sj < Ny, else 0.
    lo=0: the CUDA path's `>=0` bounds (backproplib.cu:94); lo=1: the CPU path's strict `>0` (netlib.cpp:344)."""
    Nx, Ny = x.shape[-2:]
    y = np.zeros_like(x)
    i0, i1 = max(0, lo + si), min(Nx, Nx + si)
    j0, j1 = max(0, lo + sj), min(Ny, Ny + sj)
    if i0 < i1 and j0 < j1:
        y[..., i0:i1, j0:j1] = x[..., i0 - si : i1 - si, j0 - sj : j1 - sj]
    return y


def valid_mask(Nx: int, Ny: int, si: int, sj: int, lo: int = 0) -> np.ndarray:
    """V(i-si, j-sj) as a 0/1 image."""
    return shifted(np.ones((Nx, Ny), F64), si, sj, lo)


# --------------------------------------------------------------------------------------------------
# forward: Conv (CPU), Conv_gpu, Pool, Portion
# --------------------------------------------------------------------------------------------------


def _conv(x, c, b, path, lo, pre_scale):
    x = np.asarray(x, F64)
    c = np.asarray(c, F64)
    dM, dD, Nk, Nl = c.shape
    oi, oj = tap_offsets(Nk, path), tap_offsets(Nl, path)
    xs = x * pre_scale
    out = np.zeros(x.shape[:-3] + (dM,) + x.shape[-2:], F64)
    for k in range(Nk):
        for l in range(Nl):
            sh = shifted(xs, int(oi[k]), int(oj[l]), lo)  # [..., dD, Nx, Ny]
            out += np.einsum("md,...dij->...mij", c[:, :, k, l], sh)
    out += np.asarray(b, F64)[:, None, None]
    return out


def conv_cpu(x, c, b):
    """netlib.cpp:318-358 Conv: strict >0 bounds, offsets from ak=(Nk-1)/2-1, no /dM, identity activation."""
    return _conv(x, c, b, "cpu", 1, 1.0)


def conv_gpu(x, c, b):
    """backproplib.cu:114-182 Conv_gpu + :70-111 conv_parallel: input pre-scaled by 1/dM in float32 on the host
    (:134), >=0 bounds, offsets from ak=((Nk-1)/2-1)/2."""
    dM = np.asarray(c).shape[0]
    xs = (np.asarray(x, F32) / F32(dM)).astype(F64)  # the host division is a float32 op
    return _conv(xs, c, b, "cuda", 0, 1.0)


def pool(x, scale: int, out_shape=None):
    """netlib.cpp:114-164 Pool.  scale>0: window max through an `int smax=0` accumulator (truncates toward zero,
    floors at 0, also for scale==1); scale<0: nearest-neighbour replicate into the caller-sized output."""
    x = np.asarray(x)
    D, Nx, Ny = x.shape[-3:]
    lead = x.shape[:-3]
    if scale > 0:
        s = scale
        oNx, oNy = (Nx // s, Ny // s) if out_shape is None else out_shape
        out = np.zeros(lead + (D, oNx, oNy), F32)
        # loop i=0..Nx-1 step s writes out[i/s]; windows clipped at the image edge (i+k<Nx)
        for a in range(0, Nx, s):
            for bb in range(0, Ny, s):
                if a // s >= oNx or bb // s >= oNy:
                    continue  # the reference would write out of bounds here; sizes used are divisible
                win = x[..., a : min(a + s, Nx), bb : min(bb + s, Ny)].astype(F64)
                # sequential `if v>smax: smax=(int)v` == trunc(max(0, max v)) because trunc is monotone
                mx = np.maximum(win.reshape(lead + (D, -1)).max(-1), 0.0)
                out[..., a // s, bb // s] = np.trunc(mx).astype(F32)
        return out
    s = -scale
    assert out_shape is not None or True
    oNx, oNy = (Nx * s, Ny * s) if out_shape is None else out_shape
    out = np.zeros(lead + (D, oNx, oNy), F32)
    ii = np.minimum(np.arange(oNx) // s, Nx - 1)
    jj = np.minimum(np.arange(oNy) // s, Ny - 1)
    out[...] = x[..., ii[:, None], jj[None, :]]
    return out


def portion(inp, hin, out, q: int):
    """netlib.cpp:292-315 Portion: centre crop by factor q."""
    Nx, Ny = inp.shape[-2:]
    dx, dy = (Nx - Nx // q) // 2, (Ny - Ny // q) // 2
    sl = (Ellipsis, slice(dx, dx + Nx // q), slice(dy, dy + Ny // q))
    return inp[sl].copy(), hin[sl].copy(), out[sl].copy()


# --------------------------------------------------------------------------------------------------
# coordinate-space training
# --------------------------------------------------------------------------------------------------


def clip10(g):
    """g / max(10,|g|)  (netlib.cpp:437; backproplib.cu:393; fft_backproplib.cu:617)."""
    g = np.asarray(g, F64)
    return g / np.maximum(10.0, np.abs(g))


def _delta_h(e, f, path, lo):
    """dh[m](u,v) = V(u,v) * sum_{d1,k1,l1} f[d1][m][k1][l1] * e[d1](u+ik1, v+il1)   
    (the e index ranges over the plain image; V applies to the hidden position (u,v))."""
    dD, dM, Nk, Nl = f.shape
    oi, oj = tap_offsets(Nk, path), tap_offsets(Nl, path)
    Nx, Ny = e.shape[-2:]
    dh = np.zeros(e.shape[:-3] + (dM, Nx, Ny), F64)
    for k1 in range(Nk):
        for l1 in range(Nl):
            sh = shifted(e, -int(oi[k1]), -int(oj[l1]), 0)  # e(u+ik1, v+il1)
            dh += np.einsum("dm,...dij->...mij", f[:, :, k1, l1], sh)
    if lo:
        dh[..., :lo, :] = 0
        dh[..., :, :lo] = 0
    return dh


def _corr(a, x, path, lo, Nk, Nl):
    """g[..., A, X, k, l] = sum_{i,j} a[A](i,j) * x[X](i-ik, j-il) * V(i-ik, j-il)."""
    oi, oj = tap_offsets(Nk, path), tap_offsets(Nl, path)
    g = np.zeros(a.shape[:-3] + (a.shape[-3], x.shape[-3], Nk, Nl), F64)
    for k in range(Nk):
        for l in range(Nl):
            sh = shifted(x, int(oi[k]), int(oj[l]), lo)
            g[..., k, l] = np.einsum("...aij,...xij->...ax", a, sh)
    return g


def _hin_strided(hin, Nx_stride_quirk: bool):
    """Quirk C2 (backproplib.cu:226,283,464,510): the dF term reads hin[m*Nx*Ny + (i-ik)*Nx + (j-il)].
    Identical to the intended read when Nx==Ny.  For Nx!=Ny the compiled reference reads other channels /
    out of bounds (UB); policy: intended stride.  This helper only exists to document that."""
    return hin


def coord_gradients_cuda(inp, out, hin, c, f, sym: bool, quirks: bool = True):
    """Raw (un-clipped) gradients of ONE frame for the CUDA coordinate path.
    sym=True : backprop_gpu_cc  (backproplib.cu:521-644; kernels :424-518) -> returns g (=gC+gF^T), gB, gP, mse
    sym=False: backprop_gpu     (backproplib.cu:291-418; kernels :186-288) -> returns gC, gF, gB, gP, mse
      with bug-compat quirks C1 (bias gradient keeps only d1=dD-1, :220), C3 ((j-ik) instead of (j-il) in the
      dF term for taps != (0,0), :283) and C4 (dDdF buffer keeps stale values where the shifted index is out of
      bounds, :225,282,335) when quirks=True.
    Norm = dD*dM*Nk*Nl*Nx*Ny (x2 when sym)  (:303, :533)."""
    inp, out, hin = (np.asarray(t, F64) for t in (inp, out, hin))
    c, f = np.asarray(c, F64), np.asarray(f, F64)
    dM, dD, Nk, Nl = c.shape
    Nx, Ny = inp.shape[-2:]
    norm = float(np.float32(dD * dM * Nk * Nl * Nx * Ny) * (2 if sym else 1))
    e = out - inp
    mse = float((e**2).sum() / norm)  # printed value (:356, :587)
    dh = _delta_h(e, f, "cuda", 0)
    gC = _corr(dh, inp, "cuda", 0, Nk, Nl) / norm  # [dM,dD,Nk,Nl]
    gP = e.sum((-1, -2)) / norm
    if sym:
        gF = _corr(e, hin, "cuda", 0, Nk, Nl) / norm  # [dD,dM,Nk,Nl]
        g = gC + np.swapaxes(gF, 0, 1)
        gB = dh.sum((-1, -2)) / norm
        return g, gB, gP, mse
    if not quirks:
        gF = _corr(e, hin, "cuda", 0, Nk, Nl) / norm
        gB = dh.sum((-1, -2)) / norm
        return gC, gF, gB, gP, mse
    # --- bug-compatible emulation of the launch sequence (m; d; k; l), backproplib.cu:363-417 ---
    oi, oj = tap_offsets(Nk, "cuda"), tap_offsets(Nl, "cuda")
    # C1: dDdB2 = (assignment) -> only d1 = dD-1 survives: sum_{i,j} e[dD-1](i,j) * sum_{k1,l1 valid} f[dD-1][m][k1][l1]
    gB = np.zeros(dM, F64)
    for k1 in range(Nk):
        for l1 in range(Nl):
            vm = valid_mask(Nx, Ny, int(oi[k1]), int(oj[l1]), 0)  # V(i-ik1, j-il1)
            gB += f[dD - 1, :, k1, l1] * (e[dD - 1] * vm).sum() / norm
    gF = np.zeros((dD, dM, Nk, Nl), F64)
    buf = np.zeros((Nx, Ny), F64)  # dDdF_d, zero-initialised once (:335), never re-zeroed (C4)
    for m in range(dM):
        for d in range(dD):
            for k in range(Nk):
                for l in range(Nl):
                    ik, il = int(oi[k]), int(oj[l])
                    vm = valid_mask(Nx, Ny, ik, il, 0).astype(bool)  # i-ik, j-il in bounds
                    if k == 0 and l == 0:
                        hs = shifted(hin[m], ik, il, 0)  # gradient_CFBP :226 (stride quirk C2 only)
                    else:
                        # C3 (:283): hin[m*NxNy + (i-ik)*Nx + (j-ik)]  -- column index uses ik.
                        # read is a flat-buffer read: (i-ik) in bounds is guaranteed by the mask, but (j-ik) may
                        # leave [0,Ny) and then aliases the neighbouring row of the flat array (square frames).
                        hs = _flat_read(hin, m, ik, ik, Nx, Ny)
                    buf[vm] = (e[d] * hs)[vm] / norm
                    gF[d, m, k, l] = buf.sum()
    return gC, gF, gB, gP, mse


def _flat_read(hin, m, si, sj, Nx, Ny):
    """hin_flat[m*Nx*Ny + (i-si)*Nx + (j-sj)] for all (i,j) (0 where the flat index leaves the buffer).
    Emulates the reference's raw pointer arithmetic (row aliasing included); uses stride Nx (quirk C2)."""
    flat = hin.reshape(-1)
    i = np.arange(Nx)[:, None]
    j = np.arange(Ny)[None, :]
    idx = m * Nx * Ny + (i - si) * Nx + (j - sj)
    ok = (idx >= 0) & (idx < flat.size)
    return np.where(ok, flat[np.clip(idx, 0, flat.size - 1)], 0.0)


def momentum_update(w, v, g, delmax, alpha):
    """v = (1-alpha)*del*clip10(g) + alpha*v ; w -= v   (backproplib.cu:392-396; del==delmax because adapt_rate
    ends with del=delmax, :34 -- quirk C5).  Returns (w, v)."""
    v = (1.0 - alpha) * delmax * clip10(g) + alpha * np.asarray(v, F64)
    return np.asarray(w, F64) - v, v


def backprop_gpu_cc(inp, out, hin, c, b, f, p, dc, db, df, dp, ddc, ddb, ddf, ddp, delmax, alpha, active=1):
    """backproplib.cu:521-644 (tied weights).  Batched inputs [B,...] -> mean of per-frame raw gradients, one
    update (this repo's batch extension; B=1 == the reference).  Returns dict of updated arrays + 'mse'."""
    inp, out, hin = (np.asarray(t, F64) for t in (inp, out, hin))
    if inp.ndim == 3:
        inp, out, hin = inp[None], out[None], hin[None]
    gs = [coord_gradients_cuda(inp[n], out[n], hin[n], c, f, True) for n in range(inp.shape[0])]
    g = np.mean([x[0] for x in gs], 0)
    gB = np.mean([x[1] for x in gs], 0)
    gP = np.mean([x[2] for x in gs], 0)
    mse = float(np.mean([x[3] for x in gs]))
    c2, dc2 = momentum_update(c, dc, g, delmax, alpha)
    b2, db2 = momentum_update(b, db, gB, delmax, alpha)
    p2, dp2 = momentum_update(p, dp, gP, delmax, alpha)
    f2 = np.swapaxes(c2, 0, 1).copy()  # f[d][m][k][l] = c[m][d][k][l]  (:622)
    return dict(c=c2, b=b2, f=f2, p=p2, dc=dc2, db=db2, df=np.asarray(df, F64), dp=dp2, ddc=g, ddb=gB,
                ddf=np.asarray(ddf, F64), ddp=gP, mse=mse)


def backprop_gpu(inp, out, hin, c, b, f, p, dc, db, df, dp, ddc, ddb, ddf, ddp, delmax, alpha, active=1,
                 quirks=True):
    """backproplib.cu:291-418 (independent c,f; Jacobi: all gradients from the f uploaded before the loop)."""
    inp, out, hin = (np.asarray(t, F64) for t in (inp, out, hin))
    if inp.ndim == 3:
        inp, out, hin = inp[None], out[None], hin[None]
    gs = [coord_gradients_cuda(inp[n], out[n], hin[n], c, f, False, quirks) for n in range(inp.shape[0])]
    gC, gF, gB, gP = (np.mean([x[i] for x in gs], 0) for i in range(4))
    mse = float(np.mean([x[4] for x in gs]))
    c2, dc2 = momentum_update(c, dc, gC, delmax, alpha)
    f2, df2 = momentum_update(f, df, gF, delmax, alpha)
    b2, db2 = momentum_update(b, db, gB, delmax, alpha)
    p2, dp2 = momentum_update(p, dp, gP, delmax, alpha)
    return dict(c=c2, b=b2, f=f2, p=p2, dc=dc2, db=db2, df=df2, dp=dp2, ddc=gC, ddb=gB, ddf=gF, ddp=gP, mse=mse)


def backprop_cpu_literal(inp, out, hin, c, b, f, p, delta):
    """netlib.cpp:361-451 backprop, literal step order (m; d; k; l) with the in-place sequential `f` update
    (:437-438): step (m,d,k,l) sees f already updated for all earlier steps.  No momentum.  One frame.
    O(dM*dD*Nk*Nl) numpy passes -> use for small shapes.  Returns dict(c,b,f,p,mse)."""
    inp, out, hin = (np.asarray(t, F64) for t in (inp, out, hin))
    c, b, f, p = (np.array(t, F64) for t in (c, b, f, p))
    dM, dD, Nk, Nl = c.shape
    Nx, Ny = inp.shape[-2:]
    norm = float(np.float32(dD * dM * Nk * Nl * Nx * Ny))
    oi, oj = tap_offsets(Nk, "cpu"), tap_offsets(Nl, "cpu")
    e = out - inp
    mse = float((e**2).sum())  # the CPU path prints the raw sum (:385)
    for m in range(dM):
        for d in range(dD):
            for k in range(Nk):
                for l in range(Nl):
                    ik, il = int(oi[k]), int(oj[l])
                    # hidden delta of channel m under the CURRENT f, strict bounds on the hidden position
                    dh = np.zeros((Nx, Ny), F64)
                    for k1 in range(Nk):
                        for l1 in range(Nl):
                            sh = shifted(e, -int(oi[k1]), -int(oj[l1]), 0)
                            dh += np.einsum("d,dij->ij", f[:, m, k1, l1], sh)
                    dh[:1, :] = 0
                    dh[:, :1] = 0
                    gC = (dh * shifted(inp[d], ik, il, 1)).sum() / norm
                    gB = dh.sum() / norm
                    gF = (e[d] * shifted(hin[m], ik, il, 1)).sum() / norm
                    gP = e[d].sum() / norm
                    c[m, d, k, l] -= delta * clip10(gC)
                    f[d, m, k, l] -= delta * clip10(gF)
                    if k == 0 and l == 0:
                        if d == 0:
                            b[m] -= delta * clip10(gB)
                        if m == 0:
                            p[d] -= delta * clip10(gP)
    return dict(c=c, b=b, f=f, p=p, mse=mse)


def cpu_ref_tensors(inp, out, hin, Nk, Nl):
    """f-independent tensors of the CPU path, one frame, un-normalised:
       R[d1,k1,l1,d,k,l] = sum_{u,v} V(u,v) e[d1](u+ik1,v+il1) in[d](u-ik,v-il) V(u-ik,v-il)
       Bm[d1,k1,l1]      = sum_{u,v} V(u,v) e[d1](u+ik1,v+il1)
       GF[d,m,k,l]       = sum_{i,j} e[d](i,j) hin[m](i-ik,j-il) V(i-ik,j-il);   GP[d] = sum e[d]"""
    inp, out, hin = (np.asarray(t, F64) for t in (inp, out, hin))
    dD = inp.shape[0]
    Nx, Ny = inp.shape[-2:]
    oi, oj = tap_offsets(Nk, "cpu"), tap_offsets(Nl, "cpu")
    e = out - inp
    V = np.zeros((Nx, Ny), F64)
    V[1:, 1:] = 1.0
    E = np.zeros((dD, Nk, Nl, Nx, Ny), F64)
    X = np.zeros((dD, Nk, Nl, Nx, Ny), F64)
    for k in range(Nk):
        for l in range(Nl):
            E[:, k, l] = shifted(e, -int(oi[k]), -int(oj[l]), 0) * V
            X[:, k, l] = shifted(inp, int(oi[k]), int(oj[l]), 1)
    R = np.einsum("aklij,bmnij->aklbmn", E, X)
    Bm = E.sum((-1, -2))
    GF = _corr(e, hin, "cpu", 1, Nk, Nl)
    GP = e.sum((-1, -2))
    return R, Bm, GF, GP, float((e**2).sum())


def backprop_cpu(inp, out, hin, c, b, f, p, delta):
    """netlib.cpp:361-451 via the parallel formulation (SURVEY A.3 / probe P5): gF, gP do not depend on f, so all
    f_new are known up front and gC(m,d,k,l) = (1/Norm) sum_s [s <lex (d,k,l) ? f_new : f_old][s,m] R[s;(d,k,l)].
    Batched [B,...]: tensors are averaged over frames, then one update."""
    inp, out, hin = (np.asarray(t, F64) for t in (inp, out, hin))
    if inp.ndim == 3:
        inp, out, hin = inp[None], out[None], hin[None]
    c, b, f, p = (np.array(t, F64) for t in (c, b, f, p))
    dM, dD, Nk, Nl = c.shape
    Nx, Ny = inp.shape[-2:]
    norm = float(np.float32(dD * dM * Nk * Nl * Nx * Ny))
    ts = [cpu_ref_tensors(inp[n], out[n], hin[n], Nk, Nl) for n in range(inp.shape[0])]
    R, Bm, GF, GP = (np.mean([t[i] for t in ts], 0) / norm for i in range(4))
    mse = float(np.mean([t[4] for t in ts]))
    S = dD * Nk * Nl
    f_old = np.transpose(f, (0, 2, 3, 1)).reshape(S, dM)  # [s=(d1,k1,l1), m]
    f_new = f_old - delta * clip10(np.transpose(GF, (0, 2, 3, 1)).reshape(S, dM))
    Rm = R.reshape(S, S)  # [s, t=(d,k,l)]
    lower = np.tril(np.ones((S, S)), -1).T  # lower[s,t] = 1 if s < t
    gC = np.einsum("sm,st->mt", f_new, Rm * lower) + np.einsum("sm,st->mt", f_old, Rm * (1 - lower))
    gB = np.einsum("sm,s->m", f_old, Bm.reshape(S))
    c -= delta * clip10(gC.reshape(dM, dD, Nk, Nl))
    b -= delta * clip10(gB)
    p -= delta * clip10(GP)
    f = np.transpose(f_new.reshape(dD, Nk, Nl, dM), (0, 3, 1, 2)).copy()
    return dict(c=c, b=b, f=f, p=p, mse=mse)


# --------------------------------------------------------------------------------------------------
# momentum (FFT) space -- fft_backproplib.cu.  Half spectra [ch][Nx][Nyr], unnormalised both ways.
# --------------------------------------------------------------------------------------------------


def r2c(x):
    """cufftExecR2C with n={Nx,Ny} (fft_backproplib.cu:773-796): unnormalised rfft2."""
    return np.fft.rfft2(np.asarray(x, F64), axes=(-2, -1))


def c2r(X, Ny: int):
    """cufftExecC2R (fft_backproplib.cu:821-829): unnormalised inverse (numpy's irfft2 * Nx*Ny)."""
    Nx = X.shape[-2]
    return np.fft.irfft2(X, s=(Nx, Ny), axes=(-2, -1)) * (Nx * Ny)


def kernel_pad(c, Nx: int, Ny: int):
    """fft_backproplib.cu:1018-1064 kernel_pad / :570-600 pad_k: img[(k-Nk/2) mod Nx][(l-Nl/2) mod Ny] = c[k][l]."""
    c = np.asarray(c)
    Nk, Nl = c.shape[-2:]
    out = np.zeros(c.shape[:-2] + (Nx, Ny), c.dtype)
    for k in range(Nk):
        for l in range(Nl):
            out[..., (k - Nk // 2) % Nx, (l - Nl // 2) % Ny] = c[..., k, l]
    return out


def kernel_shrink(img, Nk: int, Nl: int):
    """fft_backproplib.cu:535-565 shrink_k / :1069-1112 kernel_invpad: inverse gather of kernel_pad."""
    img = np.asarray(img)
    Nx, Ny = img.shape[-2:]
    out = np.zeros(img.shape[:-2] + (Nk, Nl), img.dtype)
    for k in range(Nk):
        for l in range(Nl):
            out[..., k, l] = img[..., (k - Nk // 2) % Nx, (l - Nl // 2) % Ny]
    return out


def kernel_spectrum(c, Nx: int, Ny: int):
    """StoreLoad_cfreq first-time branch (fft_backproplib.cu:1148-1157): R2C(kernel_pad(c))."""
    return r2c(kernel_pad(np.asarray(c, F64), Nx, Ny))


def cfreq_to_wire(C):
    """store_cfreq / copy_out (fft_backproplib.cu:1117-1127, 246-262): interleaved (re,im) float32."""
    C = np.asarray(C)
    w = np.empty(C.shape + (2,), F32)
    w[..., 0] = C.real
    w[..., 1] = C.imag
    return w.reshape(-1)


def wire_to_cfreq(w, dM, dD, Nx, Nyr):
    """load_cfreq / copy_in (fft_backproplib.cu:1131-1141, 267-282)."""
    w = np.asarray(w, F64).reshape(dM, dD, Nx, Nyr, 2)
    return w[..., 0] + 1j * w[..., 1]


def resize_spectrum(X, Nx: int, Ny: int, scale: int):
    """pool_fft + resize (fft_backproplib.cu:975-1002, 87-157): spectral pooling. scale>1 crops, scale<0 zero-embeds;
    no amplitude rescale.  X: [..., ch, Nx, Nyr].  Returns (Y, Nxs, Nys)."""
    if scale == 1:
        return X, Nx, Ny
    l = float(scale) if scale > 0 else -1.0 / float(scale)
    Nxs, Nys = int(Nx / l), int(Ny / l)
    Nyr, Nyrs = Ny // 2 + 1, Nys // 2 + 1
    Y = np.zeros(X.shape[:-2] + (Nxs, Nyrs), X.dtype)
    for i in range(Nxs):
        if Nxs <= Nx:
            src = i if i < Nxs // 2 else (Nx // 2 if i == Nxs // 2 else i + Nx - Nxs)
            Y[..., i, : Nyrs - 1] = X[..., src, : Nyrs - 1]
            Y[..., i, Nyrs - 1] = X[..., src, Nyr - 1]
        else:
            if i < Nx // 2:
                src = i
            elif i > Nxs - Nx // 2:
                src = i - Nxs + Nx
            elif i == Nxs // 2:
                src = Nx // 2
            else:
                continue
            Y[..., i, : Nyr - 1] = X[..., src, : Nyr - 1]
            Y[..., i, Nyrs - 1] = X[..., src, Nyr - 1]
    return Y, Nxs, Nys


def conv_k(X, C, b, Nx, Ny):
    """conv_k (fft_backproplib.cu:162-189): per bin out[m] = sum_d (in[d]/dM)*c[m][d]; + b[m]*Nx*Ny at DC."""
    dM = C.shape[0]
    out = np.einsum("mdij,...dij->...mij", C, X / dM)
    out[..., 0, 0] += np.asarray(b, F64) * (Nx * Ny)
    return out


def mse_fft(Xt, O, dM, Nx, Ny):
    """mse_fft + calc_mse (fft_backproplib.cu:1178-1192, 480-498): Hermitian-weighted sum |Xt-O|^2."""
    dD = Xt.shape[-3]
    Nyr = Ny // 2 + 1
    n = np.full(Nyr, float(dD * Nx * Ny))
    n[1 : Nyr - 1] /= 2
    v = (np.abs(Xt - O) ** 2 / n).sum((-1, -2, -3))
    return v / (2 * dM * Nx * Ny)


def gradient_k_io(X, Xt, O, C, Fq, b, Nx, Ny):
    """gradient_k_io (fft_backproplib.cu:395-475).  X: input spectrum, Xt: expected output, O: autoencoder output;
    C[m][d], Fq[d][m] kernel spectra.  Returns dC[m][d], dF[d][m] (spectra), db[m], dp[d].
    Includes quirk F1: H-hat is built WITHOUT the /dM that conv_k applies (:428-429 vs :176-177)."""
    dM, dD = C.shape[:2]
    norm = float(Nx * Ny)
    Norm = norm * 2 * dM * dD * Nx * Ny
    E = O - Xt  # [dD, Nx, Nyr]
    G = np.einsum("dij,dmij->mij", E, np.conj(Fq))  # sum_d1 E_d1 conj(F_d1,m)
    dC = G[:, None] * np.conj(X)[None, :] / Norm
    Hh = np.einsum("mdij,dij->mij", C, X)
    Hh[:, 0, 0] += np.asarray(b, F64) * norm
    dF = E[:, None] * np.conj(Hh)[None, :] / Norm
    db = G[:, 0, 0].real * norm / Norm
    dp = E[:, 0, 0].real * norm / Norm
    return dC, dF, db, dp


def gradient_diff(c, f, b, p):
    """gradient_diff (fft_backproplib.cu:709-753): kernel-diversity gradient; pairs need m1!=m AND d1!=d."""
    c, f, b, p = (np.asarray(t, F64) for t in (c, f, b, p))
    dM, dD = c.shape[:2]
    cd = np.zeros_like(c)
    fd = np.zeros_like(f)
    with np.errstate(divide="ignore", invalid="ignore"):
        for m in range(dM):
            for d in range(dD):
                for m1 in range(dM):
                    for d1 in range(dD):
                        if m1 != m and d1 != d:
                            dc_ = c[m, d] - c[m1, d1]
                            df_ = f[d, m] - f[d1, m1]
                            cd[m, d] += dc_ / (dc_**2).sum()
                            fd[d, m] += df_ / (df_**2).sum()
        bd = np.array([sum(1.0 / (b[m] - b[m1]) for m1 in range(dM) if m1 != m) for m in range(dM)], F64)
        pd = np.array([sum(1.0 / (p[d] - p[d1]) for d1 in range(dD) if d1 != d) for d in range(dD)], F64)
    return cd, fd, bd, pd


def gradient_diff_blocked(c, f, b, p, rows=1024):
    """Same sums as gradient_diff for the kernel counts of BASELINE config 4 (2 048 .. 32 768 kernels per tensor), where the
    literal loop above would take hours: sum_b w_ab (x_a - x_b) = x_a sum_b w_ab - sum_b w_ab x_b with
    w_ab = [m_a != m_b and d_a != d_b] / |x_a - x_b|^2 and |x_a - x_b|^2 = |x_a|^2 + |x_b|^2 - 2 x_a.x_b, `rows` kernels at
    a time (fp64: the cancellation costs ~1e-11 relative even for kernels 1e-2 apart).  The loop is the restatement;
    tests/test_oracle_cpu.py pins this form against it."""
    c, f, b, p = (np.asarray(t, F64) for t in (c, f, b, p))
    dM, dD = c.shape[:2]
    T = c.shape[2] * c.shape[3]
    xc = c.reshape(dM * dD, T)                              # row a = m * dD + d
    xf = np.swapaxes(f, 0, 1).reshape(dM * dD, T)           # f[d][m] put on the same row index a
    am, ad = np.divmod(np.arange(dM * dD), dD)
    oc, of = np.zeros_like(xc), np.zeros_like(xf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for x, o in ((xc, oc), (xf, of)):
            n2 = (x * x).sum(1)
            for a0 in range(0, dM * dD, rows):
                sl = slice(a0, min(a0 + rows, dM * dD))
                ok = (am[sl, None] != am[None, :]) & (ad[sl, None] != ad[None, :])
                d2 = n2[sl, None] + n2[None, :] - 2.0 * (x[sl] @ x.T)
                w = np.where(ok, 1.0 / d2, 0.0)
                o[sl] = x[sl] * w.sum(1)[:, None] - w @ x
        bd = np.array([sum(1.0 / (b[m] - b[m1]) for m1 in range(dM) if m1 != m) for m in range(dM)], F64)
        pd = np.array([sum(1.0 / (p[d] - p[d1]) for d1 in range(dD) if d1 != d) for d in range(dD)], F64)
    return oc.reshape(c.shape), np.swapaxes(of.reshape(dM, dD, *f.shape[2:]), 0, 1).copy(), bd, pd


def backprop_fft(inp, expout, out, c, f, b, p, del0, maxdiff=0, n_iter=100, cfreq=None, ffreq=None,
                 return_trace=False):
    """backprop_fft (fft_backproplib.cu:1381-1511).  Batched [B,...]: raw kernel-space gradients averaged over
    frames before the (non-linear) clip; B=1 == the reference.  cfreq/ffreq: cached spectra (complex arrays); when
    None they are derived from c,f (what StoreLoad_cfreq would have cached).  Returns dict(c,f,b,p,cfreq,ffreq,mse[])."""
    inp, expout, out = (np.asarray(t, F64) for t in (inp, expout, out))
    if inp.ndim == 3:
        inp, expout, out = inp[None], expout[None], out[None]
    c, f, b, p = (np.array(t, F64) for t in (c, f, b, p))
    dM, dD, Nk, Nl = c.shape
    Nx, Ny = inp.shape[-2:]
    B = inp.shape[0]
    X, Xt, O = r2c(inp), r2c(expout), r2c(out)
    C = kernel_spectrum(c, Nx, Ny) if cfreq is None else np.array(cfreq, np.complex128)
    Fq = kernel_spectrum(f, Nx, Ny) if ffreq is None else np.array(ffreq, np.complex128)
    Dc, Df, Db, Dp = np.zeros_like(c), np.zeros_like(f), np.zeros_like(b), np.zeros_like(p)
    alpha = 0.9  # hard-coded (:608, :660)
    delta = float(np.float32(0.1) * np.float32(del0))  # (:1445)
    mses = [float(np.mean(mse_fft(Xt, O, dM, Nx, Ny)))]
    for _ in range(n_iter):
        acc = [0, 0, 0, 0]
        for n in range(B):
            dC, dF, db, dp = gradient_k_io(X[n], Xt[n], O[n], C, Fq, b, Nx, Ny)
            dck = kernel_shrink(c2r(dC, Ny), Nk, Nl)  # (:1219-1226)
            dfk = kernel_shrink(c2r(dF, Ny), Nk, Nl)
            for i, t in enumerate((dck, dfk, db, dp)):
                acc[i] = acc[i] + t / B
        dck, dfk, db, dp = acc
        if maxdiff:
            cd, fd, bd, pd = (gradient_diff if dM * dD <= 512 else gradient_diff_blocked)(c, f, b, p)  # (:1237) w0=1, w1=10 (:1252)
            dck, dfk, db, dp = dck - 10 * cd, dfk - 10 * fd, db - 10 * bd, dp - 10 * pd
        c, Dc = momentum_update(c, Dc, dck, delta, alpha)  # backprop_d / backprop_double (:605-704)
        f, Df = momentum_update(f, Df, dfk, delta, alpha)
        b, Db = momentum_update(b, Db, db, delta, alpha)
        p, Dp = momentum_update(p, Dp, dp, delta, alpha)
        C = kernel_spectrum(c, Nx, Ny)  # pad_k + R2C (:1274-1282)
        Fq = kernel_spectrum(f, Nx, Ny)
        H = conv_k(X, C, b, Nx, Ny)  # (:1460)
        O = conv_k(H, Fq, p, Nx, Ny)  # (:1461)  (divides by dD, bias p)
        mses.append(float(np.mean(mse_fft(Xt, O, dM, Nx, Ny))))
    # export_cfreq (:1487-1488): c,f re-derived from the spectra (C2R/(NxNy) + kernel_invpad) == c,f up to round-off
    c_out = kernel_shrink(c2r(C, Ny) / (Nx * Ny), Nk, Nl)
    f_out = kernel_shrink(c2r(Fq, Ny) / (Nx * Ny), Nk, Nl)
    res = dict(c=c_out, f=f_out, b=b, p=p, cfreq=C, ffreq=Fq, mse=mses, O=O)
    return res


def autoenc_fft(x, net_c, net_b, scale, net_cfreq=None, fft_l=1):
    """autoenc_fft (fft_backproplib.cu:1331-1376): full-stack forward in frequency space.
    x: [dD,Nx,Ny] (or batched).  Returns list of layers (2*n_conv+1 entries; entries other than first/last are
    None when fft_l==0) and the list of kernel spectra used (what net_cfreq caches)."""
    x = np.asarray(x, F64)
    Nx, Ny = x.shape[-2:]
    n_conv = len(net_c)
    layers = [x]
    freq = r2c(x)
    spectra = []
    for n in range(n_conv):
        c = np.asarray(net_c[n], F64)
        if n < n_conv // 2:
            freq, Nx, Ny = resize_spectrum(freq, Nx, Ny, scale[n])
            layers.append(c2r(freq, Ny) / (Nx * Ny) if fft_l else None)
        C = kernel_spectrum(c, Nx, Ny) if net_cfreq is None else np.asarray(net_cfreq[n])
        spectra.append(C)
        freq = conv_k(freq, C, net_b[n], Nx, Ny)
        layers.append(c2r(freq, Ny) / (Nx * Ny) if fft_l else None)
        if n >= n_conv // 2:
            freq, Nx, Ny = resize_spectrum(freq, Nx, Ny, scale[n])
            layers.append(c2r(freq, Ny) / (Nx * Ny) if fft_l else None)
    if not fft_l:
        layers[-1] = c2r(freq, Ny) / (Nx * Ny)
    return layers, spectra


# --------------------------------------------------------------------------------------------------
# synthetic frames: counter-hash pixels 0..255, identical on CPU, every GPU and every rank
# --------------------------------------------------------------------------------------------------


def synth_frames(seed: int, B: int, D: int, Nx: int, Ny: int, b0: int = 0) -> np.ndarray:
    """frame b, channel d, pixel (i,j): float(splitmix64(seed, linear index) & 255).  This is this repo's own
    definition (the reference has no synthetic generator); the CUDA twin is aefft_synth_frames."""
    idx = (np.arange(b0, b0 + B, dtype=np.uint64)[:, None, None, None] * np.uint64(D)
           + np.arange(D, dtype=np.uint64)[None, :, None, None])
    idx = (idx * np.uint64(Nx) + np.arange(Nx, dtype=np.uint64)[None, None, :, None]) * np.uint64(Ny) \
        + np.arange(Ny, dtype=np.uint64)[None, None, None, :]
    with np.errstate(over="ignore"):
        z = idx + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z & np.uint64(255)).astype(F32)


def rel_l2(a, b) -> float:
    a, b = np.asarray(a), np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else F64  # never drop an imaginary part silently
    a, b = a.astype(dt).ravel(), b.astype(dt).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (d if d > 0 else 1.0))
