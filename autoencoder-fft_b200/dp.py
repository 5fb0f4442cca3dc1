"""Data-parallel plumbing shared by bench.py and the gloo tests: which frames a rank owns and the one collective of a
training step.  The reference has no multi-GPU path; this is the engine's extension (DESIGN.md "Multi-GPU"):
every rank computes the RAW (un-normalised, un-clipped) gradient block of its own frames, ONE all-reduce(sum) of that
small fp32 block per layer pair, then the identical clip + momentum update on every rank (weights stay replicated, no
broadcast).  The reduction must precede the clip g/max(10,|g|), which is non-linear."""
from __future__ import annotations


def frame_range(rank: int, world: int, batch_per_rank: int):
    """(first global frame index, count) of `rank` -- weak scaling: every rank owns batch_per_rank frames."""
    return rank * batch_per_rank, batch_per_rank


def allreduce_gradient_block(gbuf, world: int):
    """Sum the raw gradient block over ranks in place (NCCL on GPUs, gloo in the CPU tests)."""
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(gbuf, op=dist.ReduceOp.SUM)
    return gbuf


def bin_slab(rank: int, world: int, Ny: int):
    """(first column, column count) of the half spectrum (Ny//2+1 columns) owned by `rank` under frequency-bin sharding;
    the same split as aefft_set_bin_shard / backprop_fft_core (csrc/fft_capi.cu)."""
    nyr = Ny // 2 + 1
    c0 = rank * nyr // world
    return c0, (rank + 1) * nyr // world - c0


def allreduce_partial_block(block, world: int):
    """Bin-sharded devices hold PARTIAL sums (over their spectrum columns) of the gradient block and of the mse: add them."""
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(block, op=dist.ReduceOp.SUM)
    return block
