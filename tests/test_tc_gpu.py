"""GPU parity tests of the tensor-core (tcgen05) coordinate-space kernels against the numpy oracle.
BF16X3 (fp32 operands split into bf16 hi+lo, three products, fp32 accumulation in TMEM) must meet the fp32 bar of
north_star (1e-4 relative L2); BF16 (single product) is the documented looser mode (1e-2)."""
import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O
from test_coord_gpu import check_weights, make_case, run_product, zeros_state

pytestmark = pytest.mark.gpu

SHAPES = [(16, 3, 5, 5, 20, 14), (16, 3, 5, 5, 70, 66), (32, 16, 5, 5, 33, 65), (64, 32, 5, 5, 40, 30), (3, 16, 5, 5, 64, 48),
          (32, 64, 5, 5, 24, 20), (8, 1, 5, 5, 64, 48), (5, 4, 3, 3, 17, 19), (20, 6, 7, 7, 31, 40), (16, 3, 5, 5, 320, 240)]


@pytest.fixture()
def tc(ctx):
    yield ctx
    ctx.set_precision(A.PRECISION_FP32)


@pytest.mark.parametrize("dims", SHAPES)
@pytest.mark.parametrize("conv", ["cuda", "cpu"])
def test_conv_fwd_bf16x3(tc, dims, conv):
    cs = make_case(21, *dims, conv=conv)
    tc.set_precision(A.PRECISION_BF16X3)
    l0 = tc.launches
    got = tc.conv_fwd(cs["inp"], cs["c"], cs["b"], A.CONV_CUDA if conv == "cuda" else A.CONV_CPU)
    assert tc.launches - l0 == 2, "expected weight_prep + conv_tc (tensor-core path), not the fp32 fallback"
    want = (O.conv_gpu if conv == "cuda" else O.conv_cpu)(cs["inp"], cs["c"], cs["b"])
    assert O.rel_l2(got, want) < 3e-5


def test_conv_fwd_bf16x3_batched(tc):
    cs = make_case(22, 32, 16, 5, 5, 48, 40, B=5)
    tc.set_precision(A.PRECISION_BF16X3)
    got = tc.conv_fwd(cs["inp"], cs["c"], cs["b"])
    for n in range(5):
        assert O.rel_l2(got[n], O.conv_gpu(cs["inp"][n], cs["c"], cs["b"])) < 3e-5


def test_conv_fwd_bf16_single_pass_looser(tc):
    cs = make_case(23, 32, 16, 5, 5, 48, 40)
    tc.set_precision(A.PRECISION_BF16)
    got = tc.conv_fwd(cs["inp"], cs["c"], cs["b"])
    r = O.rel_l2(got, O.conv_gpu(cs["inp"], cs["c"], cs["b"]))
    assert 1e-5 < r < 1e-2  # genuinely the low-precision path, inside its stated tolerance


@pytest.mark.parametrize("dims", [(16, 3, 5, 5, 48, 48), (32, 16, 5, 5, 40, 72), (64, 32, 5, 5, 24, 20), (8, 5, 3, 3, 33, 17)])
@pytest.mark.parametrize("mode", ["sym", "cuda"])
def test_backprop_bf16x3(tc, dims, mode):
    cs = make_case(24, *dims, B=2)
    st = zeros_state(cs)
    tc.set_precision(A.PRECISION_BF16X3)
    if mode == "sym":
        cs["f"] = np.ascontiguousarray(np.swapaxes(cs["c"], 0, 1))
        got = run_product(tc, A.MODE_CUDA_REF_SYM, cs, st)
        want = O.backprop_gpu_cc(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], **st, delmax=0.2, alpha=0.9)
    else:
        got = run_product(tc, A.MODE_CUDA_REF, cs, st, quirks=0)
        want = O.backprop_gpu(cs["inp"], cs["out"], cs["hin"], cs["c"], cs["b"], cs["f"], cs["p"], **st, delmax=0.2, alpha=0.9,
                              quirks=False)
    # weights and biases at 1e-4 as everywhere; the UPDATE (delta) of the kernels at 1e-3.  The bias updates are sums of
    # e / dh with near-total cancellation (|sum| << sum|.|), so their delta is only meaningful to ~1e-2 in fp32 inputs
    check_weights(got, want, cs, keys="cf")
    check_weights(got, want, cs, keys="bp", tol_dw=2e-2)
