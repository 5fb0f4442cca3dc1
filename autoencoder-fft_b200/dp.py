"""Host-side sharding plan of the multi-GPU paths (one process per GPU), shared by bench.py and the CPU (gloo) tests.

The reference has no multi-GPU path; the engine adds two splits (DESIGN.md "Multi-GPU"):
  * data-parallel frames: rank r owns frames [r*B, (r+1)*B) of the global batch; the raw (un-normalised, un-clipped)
    gradient blocks of ALL layer pairs form one fused fp32 buffer that is all-reduced (sum) ONCE per step, then every rank
    applies the identical clip + momentum update (weights stay replicated, no broadcast).  The reduction must precede the
    clip g/max(10,|g|), which is non-linear.  The collective itself is issued by the engine (csrc/comm.cu, NCCL on the
    ctx stream); the functions below only say WHO owns WHAT, so that the plan can be tested without a GPU.
  * frequency-bin sharding: rank r owns the spectrum columns bin_slab(r) of every image, and rows row_slab(r) of every
    frame for the row pass of the slab-decomposed 2-D transform (rows -> all-to-all -> columns).
"""
from __future__ import annotations


def frame_range(rank: int, world: int, batch_per_rank: int):
    """(first global frame index, count) of `rank` -- weak scaling: every rank owns batch_per_rank frames."""
    return rank * batch_per_rank, batch_per_rank


def gbuf_len(mode: int, dD: int, dM: int, Nk: int, Nl: int) -> int:
    """Length (floats) of one pair's raw gradient block; mirrors aefft_coord_gbuf_len (csrc/capi.cu gbuf_len).
    mode 0 = CPU_REF: [R (dD*T)^2 | Bm dD*T | GF dD*dM*T | GP dD | SQ 1]; modes 1, 2 = CUDA_REF(_SYM):
    [GC dM*dD*T | GF dD*dM*T | GB dM | GP dD | SQ 1]."""
    T = Nk * Nl
    nC = dM * dD * T
    if mode == 0:
        S = dD * T
        return S * S + S + nC + dD + 1
    return 2 * nC + dM + dD + 1


def fused_block_layout(pairs, mode: int):
    """pairs: [(dD, dM, Nk, Nl)] in pair order.  Returns (offsets, total) of the fused gradient buffer
    [pair 0 | pair 1 | ...] -- the same layout as aefft_net_fused_layout."""
    offs, total = [], 0
    for dD, dM, Nk, Nl in pairs:
        offs.append(total)
        total += gbuf_len(mode, dD, dM, Nk, Nl)
    return offs, total


def bin_slab(rank: int, world: int, Ny: int):
    """(first column, column count) of the half spectrum (Ny//2+1 columns) owned by `rank` under frequency-bin sharding;
    the same split as aefft_set_bin_shard / backprop_fft_core (csrc/fft_capi.cu)."""
    nyr = Ny // 2 + 1
    c0 = rank * nyr // world
    return c0, (rank + 1) * nyr // world - c0


def row_slab(rank: int, world: int, Nx: int):
    """(first row, row count) of every frame that `rank` row-transforms in the slab-decomposed 2-D R2C."""
    r0 = rank * Nx // world
    return r0, (rank + 1) * Nx // world - r0


def slab_exchange_plan(rank: int, world: int, images_per_rank: int, Nx: int, Ny: int):
    """Counts / offsets (in floats; complex = 2 floats) of the all-to-all that turns frame-sharded FULL half spectra
    [images_per_rank][Nx][Ny//2+1] into this rank's column slab of ALL ranks' images -- the same arithmetic as the engine
    (csrc/net_fft.cu, bin-sharded training step): the block sent to rank r is the slab bin_slab(r) of every local image,
    [images_per_rank][Nx][ncols_r]; the blocks received, ordered by source rank, form [world * images_per_rank][Nx][ncols_me].
    Returns (send_counts, send_offsets, recv_counts, recv_offsets)."""
    _, my_nc = bin_slab(rank, world, Ny)
    scount, soff, rcount, roff, so = [], [], [], [], 0
    for r in range(world):
        _, nc = bin_slab(r, world, Ny)
        scount.append(2 * images_per_rank * Nx * nc)
        soff.append(so)
        so += scount[-1]
        rcount.append(2 * images_per_rank * Nx * my_nc)
        roff.append(r * rcount[-1])
    return scount, soff, rcount, roff
