"""GPU parity tests of the row-streaming tcgen05 convolution (conv_rs.cu) against the numpy oracle (Conv_gpu semantics,
backproplib.cu:70-182): every geometry class of the kernel -- one / several column strips, two narrow frames sharing
an M-block (with an odd number of them), band splits, output-channel jobs, channel padding, C <= 8 tap packing, several
K stages -- plus the data-gradient use (transposed window, e = out - in fused) through the gradient block.
BF16X3 must meet 3e-5 relative L2 on conv outputs (fp32 bar 1e-4)."""
import numpy as np
import pytest

import aefft_ctypes as A
import oracle_np as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc(ctx):
    ctx.set_precision(A.PRECISION_BF16X3)
    yield ctx
    ctx.set_precision(A.PRECISION_FP32)


# (dM, dD, Nk, Nl, Nx, Ny, B)
SHAPES = [
    (16, 3, 5, 5, 37, 44, 3),      # C <= 8 tap packing, one 64-pixel strip, odd number of narrow frames (G = 2)
    (3, 16, 5, 5, 40, 60, 2),      # 3 outputs padded to 16, exactly full 64-pixel strips
    (32, 16, 5, 5, 33, 128, 2),    # two 124-pixel strips, the second nearly empty
    (32, 16, 5, 5, 50, 252, 1),    # three strips
    (64, 32, 5, 5, 24, 20, 5),     # two K stages, output channels split into jobs
    (32, 64, 5, 5, 20, 24, 3),     # four K stages (single-slot staging)
    (20, 6, 7, 7, 31, 40, 2),      # 7x7 window, odd channel counts
    (5, 4, 3, 3, 17, 16, 4),       # 3x3 window
    (8, 1, 5, 5, 64, 48, 1),       # BASELINE config 1 shape class (D = 1)
    (16, 3, 5, 5, 9, 8, 1),        # frame smaller than the window halo
    (16, 3, 5, 5, 320, 240, 2),    # config 2, pair 0
    (3, 16, 5, 5, 320, 240, 2),    # config 2, pair 0 decoder: (window column, output) packed N, column sums in the epilogue
    (2, 32, 7, 7, 45, 136, 3),     # the same form with two K stages, 7x7 window (14 of 16 columns), two strips
    (32, 16, 5, 5, 160, 120, 2),   # config 2, pair 1
]


@pytest.mark.parametrize("dims", SHAPES)
def test_conv_rs_forward_vs_oracle(tc, dims):
    dM, dD, Nk, Nl, Nx, Ny, B = dims
    rng = np.random.default_rng(41)
    x = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    tc.profile_enable(True)
    got = tc.conv_fwd(x, c, b)
    names = [r["name"] for r in tc.profile_records()]
    tc.profile_enable(False)
    assert names == ["conv_fwd_rs"], f"row-streaming kernel not used: {names}"
    for n in range(B):
        assert O.rel_l2(got[n], O.conv_gpu(x[n], c, b)) < 3e-5, n


FORMS = [{"AEFFT_RS_ONE": "1"}, {"AEFFT_RS_TWO": "1"}, {"AEFFT_RS_NO_TAPPACK": "1"}, {"AEFFT_RS_STACK2": "1"},
         {"AEFFT_RS_NO_NPACK": "1"}, {"AEFFT_RS_ONE": "1", "AEFFT_RS_NO_NPACK": "1"}, {"AEFFT_RS_RU": "1"},
         {"AEFFT_RS_ONE": "1", "AEFFT_RS_STACK2": "1"}]


@pytest.mark.parametrize("env", FORMS, ids=lambda e: "+".join(sorted(e)))
@pytest.mark.parametrize("dims", [SHAPES[0], SHAPES[1], SHAPES[2], SHAPES[4], SHAPES[10]])
def test_conv_rs_forced_forms_vs_oracle(tc, dims, env, monkeypatch):
    """The kernel picks CTAs per SM, K packing and the B operand form per layer; every alternative form is forced here."""
    for k in ("AEFFT_RS_ONE", "AEFFT_RS_TWO", "AEFFT_RS_NO_TAPPACK", "AEFFT_RS_STACK2", "AEFFT_RS_NO_NPACK", "AEFFT_RS_RU"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    dM, dD, Nk, Nl, Nx, Ny, B = dims
    rng = np.random.default_rng(44)
    x = np.floor(rng.random((B, dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    tc.profile_enable(True)
    got = tc.conv_fwd(x, c, b)
    names = [r["name"] for r in tc.profile_records()]
    tc.profile_enable(False)
    assert names == ["conv_fwd_rs"], f"row-streaming kernel not used: {names}"
    for n in range(B):
        assert O.rel_l2(got[n], O.conv_gpu(x[n], c, b)) < 3e-5, n


def test_conv_rs_is_deterministic(tc):
    dM, dD, Nk, Nl, Nx, Ny, B = 32, 16, 5, 5, 64, 124, 4
    rng = np.random.default_rng(42)
    x = rng.standard_normal((B, dD, Nx, Ny)).astype(np.float32) * 50
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    b = np.zeros(dM, np.float32)
    a0 = tc.conv_fwd(x, c, b)
    a1 = tc.conv_fwd(x, c, b)
    assert np.array_equal(a0, a1)


def test_unaligned_rows_fall_back(tc):
    dM, dD, Nk, Nl, Nx, Ny = 16, 3, 5, 5, 20, 14  # row length not a multiple of 4: outside the TMA envelope
    rng = np.random.default_rng(43)
    x = np.floor(rng.random((dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    b = np.zeros(dM, np.float32)
    tc.profile_enable(True)
    got = tc.conv_fwd(x, c, b)
    names = [r["name"] for r in tc.profile_records()]
    tc.profile_enable(False)
    assert "conv_fwd_rs" not in names
    assert O.rel_l2(got, O.conv_gpu(x, c, b)) < 3e-5
