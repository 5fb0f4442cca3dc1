"""GPU tests of the tensor-core momentum-space contraction (csrc/spec_tc.cu): the batched-over-bins real GEMM on
tcgen05 (kind::tf32, 3xTF32 split) against numpy fp64, for the operand forms the five contractions of a backprop_fft
iteration use (K-major x K-major: conv_k; K-major x MN-major: G; MN-major x MN-major + complex-pair epilogue: dC, dF),
and the whole path against the fp64 oracle and against the CUDA-core kernels it replaces.  Tolerance: 1e-5 relative L2
per GEMM (measured ~1e-6), 1e-4 on trained weights (north_star)."""
import ctypes as C
import os

import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O

pytestmark = pytest.mark.gpu


def bin_gemm(ctx, a, a_mn, b, b_mn, M, N, K, outer=0, conj_out=0, scale=1.0):
    S = a.shape[0]
    da, db = ctx.to_device(a), ctx.to_device(b)
    out = A.DevBuf(ctx, (S, M // 2, N // 2, 2) if outer else (S, M, N))
    A._chk(A.lib().aefft_spec_bin_gemm(ctx.h, C.c_int64(S), A._ptr(da), a.shape[1], a.shape[2], a_mn, A._ptr(db), b.shape[1],
                                       b.shape[2], b_mn, M, N, K, outer, conj_out, C.c_float(scale), A._ptr(out)))
    ctx.sync()
    res = out.numpy()
    for t in (da, db, out):
        t.free()
    return res


@pytest.mark.parametrize("S,M,N,K", [(5, 128, 64, 32), (300, 128, 128, 64), (3, 8, 32, 16), (7, 200, 16, 96), (2, 128, 512, 32),
                                     (149, 24, 256, 256)])
def test_bin_gemm_kmajor(ctx, S, M, N, K):
    """D_w = A_w B_w^T with both operands K-major (conv_k form): rows beyond M / K beyond the tensor are TMA zero fill."""
    rng = np.random.default_rng(S + M + N + K)
    a = (rng.standard_normal((S, M, K)) * 100).astype(np.float32)
    b = rng.standard_normal((S, N, K)).astype(np.float32)
    got = bin_gemm(ctx, a, 0, b, 0, M, N, K, scale=0.5)
    want = 0.5 * np.einsum("smk,snk->smn", a.astype(np.float64), b.astype(np.float64))
    assert O.rel_l2(got, want) < 1e-5


@pytest.mark.parametrize("S,M,N,K", [(4, 128, 64, 32), (200, 128, 128, 64), (3, 16, 64, 16), (2, 96, 512, 64)])
def test_bin_gemm_b_mnmajor(ctx, S, M, N, K):
    """A K-major, B MN-major ([K][N] in memory): the G = E conj(F) form, which reads the O-GEMM's weight block transposed."""
    rng = np.random.default_rng(S * 3 + M + N + K)
    a = rng.standard_normal((S, M, K)).astype(np.float32)
    b = (rng.standard_normal((S, K, N)) * 10).astype(np.float32)
    got = bin_gemm(ctx, a, 0, b, 1, M, N, K)
    want = np.einsum("smk,skn->smn", a.astype(np.float64), b.astype(np.float64))
    assert O.rel_l2(got, want) < 1e-5


@pytest.mark.parametrize("S,P,Q,B,conj", [(4, 64, 32, 128, 0), (150, 32, 16, 128, 1), (3, 8, 8, 5, 0), (2, 128, 64, 40, 1),
                                          (5, 16, 8, 300, 0)])
def test_bin_gemm_outer_complex(ctx, S, P, Q, B, conj):
    """Frame-reduced outer product of interleaved complex operands: out[s][p][q] = sum_b P[b][p] conj(Q[b][q]) (or its
    conjugate), both operands MN-major (frames are the reduction rows)."""
    rng = np.random.default_rng(S + P + Q + B)
    pc = rng.standard_normal((S, B, P)) + 1j * rng.standard_normal((S, B, P))
    qc = rng.standard_normal((S, B, Q)) + 1j * rng.standard_normal((S, B, Q))
    inter = lambda z: np.ascontiguousarray(np.stack([z.real, z.imag], -1).reshape(z.shape[0], z.shape[1], -1).astype(np.float32))
    pa, qa = inter(pc), inter(qc)
    got = bin_gemm(ctx, pa, 1, qa, 1, 2 * P, 2 * Q, B, outer=1, conj_out=conj, scale=0.25)
    p64 = pa.astype(np.float64).reshape(S, B, P, 2)
    q64 = qa.astype(np.float64).reshape(S, B, Q, 2)
    want = 0.25 * np.einsum("sbp,sbq->spq", p64[..., 0] + 1j * p64[..., 1], np.conj(q64[..., 0] + 1j * q64[..., 1]))
    if conj:
        want = np.conj(want)
    assert O.rel_l2(got, np.stack([want.real, want.imag], -1)) < 1e-5


def _fft_case(seed, dM, dD, Nk, Nl, Nx, Ny, B):
    from test_fft_gpu import fft_case

    return fft_case(seed, dM, dD, Nk, Nl, Nx, Ny, B=B, wscale=0.1)


@pytest.mark.parametrize("dims,B", [((32, 16, 5, 5, 32, 32), 16), ((64, 32, 5, 5, 16, 16), 16), ((16, 8, 3, 3, 16, 32), 17),
                                    ((8, 8, 5, 5, 16, 16), 130)])
def test_backprop_fft_tc_vs_cuda_core_and_oracle(ctx, dims, B):
    """The tensor path against the CUDA-core path it replaces (AEFFT_NO_SPEC_TC=1) and against the fp64 oracle."""
    cs = _fft_case(51, *dims, B=B)
    runs = {}
    # tc: adjoint + two outer products (AEFFT_NO_GRAM); gram: both gradient spectra from Mg = sum_b E conj(X)
    # (gram_grad_kernel, forced: by default the engine picks it where it is cheaper); cc: the CUDA-core path
    # loop: the whole iteration loop on per-bin Gram matrices (spec_gram.cu), forced where the cost model would not pick it
    for tag, envs in (("tc", ("AEFFT_NO_GRAM", "AEFFT_NO_GRAM_LOOP")), ("gram", ("AEFFT_FORCE_GRAM", "AEFFT_NO_GRAM_LOOP")),
                      ("loop", ("AEFFT_FORCE_GRAM_LOOP",)), ("cc", ("AEFFT_NO_SPEC_TC", "AEFFT_NO_GRAM_LOOP"))):
        for env in envs:
            os.environ[env] = "1"
        try:
            w = {k: cs[k].copy() for k in "cfbp"}
            ctx.profile_enable(True)
            trace = ctx.backprop_fft(cs["inp"], cs["inp"], cs["out"], w["c"], w["f"], w["b"], w["p"], 0.2, 0, 3)
            names = {r["name"] for r in ctx.profile_records()}
            ctx.profile_enable(False)
        finally:
            for env in envs:
                os.environ.pop(env, None)
        runs[tag] = (w, trace, names)
    assert {"spec_contract_tc", "spec_outer_tc"} <= runs["tc"][2], runs["tc"][2]
    assert "spec_gram_grad" not in runs["tc"][2] and "spec_gram_grad" in runs["gram"][2], runs["gram"][2]
    assert {"spec_gram_stats", "spec_gram_iter"} <= runs["loop"][2] and "spec_outer_tc" not in runs["loop"][2], runs["loop"][2]
    assert not any(n.endswith("_tc") for n in runs["cc"][2]), runs["cc"][2]
    for tag in ("tc", "gram", "loop"):
        for k in "cfbp":
            assert O.rel_l2(runs[tag][0][k], runs["cc"][0][k]) < 2e-5, (tag, k)
        assert np.allclose(runs[tag][1], runs["cc"][1], rtol=1e-4), tag
    if B <= 17:
        want = O.backprop_fft(cs["inp"], cs["inp"], cs["out"], cs["c"], cs["f"], cs["b"], cs["p"], 0.2, 0, 3)
        for tag in ("tc", "gram", "loop"):
            assert np.allclose(runs[tag][1], want["mse"], rtol=2e-4), tag
            for k in "cfbp":
                assert O.rel_l2(runs[tag][0][k], want[k]) < 1e-4, (tag, k)
                assert O.rel_l2(runs[tag][0][k].astype(np.float64) - cs[k], want[k] - cs[k]) < 2e-3, (tag, k)
