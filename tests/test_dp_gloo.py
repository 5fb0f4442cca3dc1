"""CPU test of the multi-rank plan (world_size 2, gloo): rank r computes the oracle's raw gradient sums of ITS frames
(dp.frame_range) for TWO layer pairs into one fused block (dp.fused_block_layout, the layout aefft_net_step all-reduces
once per step), ONE all-reduce(sum), every rank applies the same clipped-momentum update -> identical weights on both
ranks, equal to the single-process full-batch step.  Also shows why the reduction must come before the clip, and that
the bin / row slabs of the sharded transform partition their axes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "autoencoder-fft_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _case():
    import oracle_np as O

    rng = np.random.default_rng(0)
    B, dD, dM, Nk, Nl, Nx, Ny = 4, 2, 3, 3, 3, 10, 8
    inp = O.synth_frames(1234, B, dD, Nx, Ny)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    p = (rng.random(dD) * 2 - 1).astype(np.float32)
    hin = O.conv_gpu(inp, c, b).astype(np.float32)
    out = O.conv_gpu(hin, f, p).astype(np.float32)
    return inp, out, hin, c, b, f, p


def _raw_sums(inp, out, hin, c, f):
    """[g | gB | gP | sq] summed over the given frames (what aefft_coord_gradients returns for CUDA_REF_SYM, in the
    combined form g = gC + gF^T), un-normalised."""
    import oracle_np as O

    acc = None
    for n in range(inp.shape[0]):
        g, gB, gP, mse = O.coord_gradients_cuda(inp[n], out[n], hin[n], c, f, True)
        v = np.concatenate([g.ravel(), gB.ravel(), gP.ravel(), [mse]])
        acc = v if acc is None else acc + v
    return acc


def _case2():
    """A second, wider pair (the fused block carries every pair of the net)."""
    import oracle_np as O

    rng = np.random.default_rng(1)
    B, dD, dM, Nk, Nl, Nx, Ny = 4, 3, 4, 3, 3, 6, 6
    inp = O.synth_frames(99, B, dD, Nx, Ny)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * 0.2).astype(np.float32)
    f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    p = (rng.random(dD) * 2 - 1).astype(np.float32)
    hin = O.conv_gpu(inp, c, b).astype(np.float32)
    out = O.conv_gpu(hin, f, p).astype(np.float32)
    return inp, out, hin, c, b, f, p


def _worker(rank, world, port, ret):
    import dp
    import oracle_np as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = [_case(), _case2()]
    shapes = [(cs[3].shape[1], cs[3].shape[0], cs[3].shape[2], cs[3].shape[3]) for cs in cases]  # (dD, dM, Nk, Nl)
    offs, total = dp.fused_block_layout(shapes, 2)
    fused = torch.zeros(total, dtype=torch.float64)
    for (inp, out, hin, c, b, f, p), off in zip(cases, offs):
        b0, n = dp.frame_range(rank, world, inp.shape[0] // world)
        sl = slice(b0, b0 + n)
        v = _raw_sums(inp[sl], out[sl], hin[sl], c, f)   # combined form [g | gB | gP | sq]
        fused[off:off + len(v)] = torch.from_numpy(v)
    dist.all_reduce(fused, op=dist.ReduceOp.SUM)          # the ONE collective of the step
    res = []
    for (inp, out, hin, c, b, f, p), off in zip(cases, offs):
        g = fused[off:].numpy() / inp.shape[0]            # mean over the GLOBAL batch
        nC = c.size
        c2, _ = O.momentum_update(c, np.zeros_like(c), g[:nC].reshape(c.shape), 0.2, 0.9)
        b2, _ = O.momentum_update(b, np.zeros_like(b), g[nC:nC + b.size], 0.2, 0.9)
        res.append((c2, b2))
    ret[rank] = res
    dist.destroy_process_group()


def test_two_rank_step_equals_full_batch():
    import oracle_np as O

    world, port = 2, 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    z = lambda a: np.zeros_like(a)
    for k, (inp, out, hin, c, b, f, p) in enumerate([_case(), _case2()]):
        want = O.backprop_gpu_cc(inp, out, hin, c, b, f, p, z(c), z(b), z(f), z(p), z(c), z(b), z(f), z(p), 0.2, 0.9)
        for r in range(world):
            c2, b2 = ret[r][k]
            assert O.rel_l2(c2, want["c"]) < 1e-12 and O.rel_l2(b2, want["b"]) < 1e-12
        assert np.array_equal(ret[0][k][0], ret[1][k][0])  # replicas stay bit-identical without any broadcast


def test_fused_layout_matches_the_engine_formula():
    """dp.gbuf_len restates aefft_coord_gbuf_len (exported by libaefft.so; a host-only function)."""
    import aefft_ctypes as A
    import dp

    for mode in (0, 1, 2):
        for dD, dM, Nk, Nl in [(3, 16, 5, 5), (16, 32, 5, 5), (32, 64, 3, 7), (1, 8, 5, 5)]:
            assert dp.gbuf_len(mode, dD, dM, Nk, Nl) == A.lib().aefft_coord_gbuf_len(mode, dD, dM, Nk, Nl)
    offs, total = dp.fused_block_layout([(3, 16, 5, 5), (16, 32, 5, 5), (32, 64, 5, 5)], 2)
    assert offs == [0, 2420, 2420 + 25649] and total == 2420 + 25649 + 102497


def test_clip_is_nonlinear_so_reduce_raw_gradients():
    import oracle_np as O

    g1, g2 = np.array([30.0, 2.0]), np.array([-10.0, 2.0])
    assert not np.allclose(O.clip10((g1 + g2) / 2), (O.clip10(g1) + O.clip10(g2)) / 2)


def test_frame_range_partitions_the_global_batch():
    import dp

    owned = [dp.frame_range(r, 4, 16) for r in range(4)]
    assert owned == [(0, 16), (16, 16), (32, 16), (48, 16)]


# ---------------------------------------------------------------------------------------------- frequency-bin sharding
def test_row_slabs_partition_the_rows():
    import dp

    for Nx in (8, 64, 2048):
        for world in (1, 2, 3, 8):
            slabs = [dp.row_slab(r, world, Nx) for r in range(world)]
            assert slabs[0][0] == 0 and sum(n for _, n in slabs) == Nx
            for (r0, n), (r1, _) in zip(slabs, slabs[1:]):
                assert r0 + n == r1


def test_bin_slabs_partition_the_half_spectrum():
    import dp

    for Ny in (16, 64, 2048):
        for world in (1, 2, 3, 8):
            slabs = [dp.bin_slab(r, world, Ny) for r in range(world)]
            assert slabs[0][0] == 0 and sum(n for _, n in slabs) == Ny // 2 + 1
            for (c0, n), (c1, _) in zip(slabs, slabs[1:]):
                assert c0 + n == c1 and n > 0


def _pruned_taps_partial(Z, Ny, Nk, Nl, c0, n):
    """Kernel-space taps from the spectrum columns [c0, c0+n) only: the pruned inverse DFT of csrc/spectral_kernels.cu
    (spectrum_to_taps): taps[k][l] = sum_{wx, wy in slab} h(wy) Re(Z[wx][wy] conj(W_Nx^(wx i_k)) conj(W_Ny^(wy j_l)))."""
    Nx = Z.shape[0]
    wx = np.arange(Nx)[:, None]
    wy = np.arange(c0, c0 + n)[None, :]
    h = np.where((wy == 0) | (wy == Ny // 2), 1.0, 2.0)
    out = np.zeros((Nk, Nl))
    for k in range(Nk):
        i = (k - Nk // 2) % Nx
        for l in range(Nl):
            j = (l - Nl // 2) % Ny
            ph = np.exp(2j * np.pi * (wx * i / Nx + wy * j / Ny))
            out[k, l] = np.sum(h * np.real(Z[:, c0:c0 + n] * ph))
    return out


def _shard_worker(rank, world, port, ret):
    import dp

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    Nx, Ny, Nk, Nl = 8, 16, 5, 5
    img = rng.standard_normal((Nx, Ny))
    Z = np.fft.rfft2(img)  # every rank holds the same frames; it keeps only its slab of columns
    c0, n = dp.bin_slab(rank, world, Ny)
    part = torch.from_numpy(_pruned_taps_partial(Z, Ny, Nk, Nl, c0, n))
    dist.all_reduce(part, op=dist.ReduceOp.SUM)  # bin-sharded devices hold PARTIAL sums of the gradient block: add them
    ret[rank] = part.numpy()
    dist.destroy_process_group()


def test_two_rank_bin_sharded_kernel_gradient_equals_full_transform():
    """sum over ranks of the slab-partial pruned inverse DFT == shrink_k(C2R(spectrum)) of the whole spectrum."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shard_worker, args=(world, port, ret), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    Nx, Ny, Nk, Nl = 8, 16, 5, 5
    img = rng.standard_normal((Nx, Ny))
    full = np.fft.irfft2(np.fft.rfft2(img), s=(Nx, Ny)) * Nx * Ny  # unnormalised C2R
    want = np.array([[full[(k - Nk // 2) % Nx, (l - Nl // 2) % Ny] for l in range(Nl)] for k in range(Nk)])
    assert np.allclose(ret[0], want, rtol=1e-10, atol=1e-9)
    assert np.array_equal(ret[0], ret[1])


def _a2a_worker(rank, world, port, ret):
    import dp

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, dD, Nx, Ny = 3, 2, 8, 12          # Ny//2+1 = 7 columns: slabs of 3 and 4
    img = B * dD
    rng = np.random.default_rng(11)
    frames = rng.standard_normal((world * B, dD, Nx, Ny))
    b0, nb = dp.frame_range(rank, world, B)
    mine = np.fft.rfft2(frames[b0:b0 + nb])                      # the data-parallel forward: FULL spectra of my frames
    scount, soff, rcount, roff = dp.slab_exchange_plan(rank, world, img, Nx, Ny)
    send = np.zeros(soff[-1] + scount[-1])
    for r in range(world):                                       # launch_spec_slab: slab r of every local image, packed
        c0, nc = dp.bin_slab(r, world, Ny)
        blk = np.ascontiguousarray(mine[..., c0:c0 + nc]).reshape(img, Nx, nc)
        send[soff[r]:soff[r] + scount[r]] = blk.view(np.float64).ravel()
    # (gloo has no all_to_all on every build: the exchange goes through an all_gather of the send buffers and of their
    # count / offset tables, and every rank then picks the block each source addressed to it -- what ncclSend/Recv delivers)
    outs = [None] * world
    pad = max(world * max(scount), 1)
    buf = torch.zeros(pad, dtype=torch.float64)
    buf[:send.size] = torch.from_numpy(send)
    meta = torch.tensor(scount + soff, dtype=torch.int64)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(bufs, buf)
    dist.all_gather(metas, meta)
    for src in range(world):
        cnt, off = int(metas[src][rank]), int(metas[src][world + rank])
        assert cnt == rcount[src]
        outs[src] = bufs[src][off:off + cnt]
    recv = np.zeros(roff[-1] + rcount[-1])
    for src in range(world):
        recv[roff[src]:roff[src] + rcount[src]] = outs[src].numpy()
    ret[rank] = recv
    dist.destroy_process_group()


def test_two_rank_slab_exchange_yields_the_frame_major_slab():
    """The all-to-all of the bin-sharded step (csrc/net_fft.cu; plan restated in dp.slab_exchange_plan): after it every rank
    holds [world * B][dD][Nx][its columns] of the GLOBAL batch, frames in global order -- exactly the slab of the global
    spectrum, with nothing but the exchange's own counts and offsets."""
    import dp

    world, port = 2, 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_a2a_worker, args=(world, port, ret), nprocs=world, join=True)
    B, dD, Nx, Ny = 3, 2, 8, 12
    rng = np.random.default_rng(11)
    full = np.fft.rfft2(rng.standard_normal((world * B, dD, Nx, Ny)))
    for rank in range(world):
        c0, nc = dp.bin_slab(rank, world, Ny)
        got = ret[rank].view(np.complex128).reshape(world * B, dD, Nx, nc)
        assert np.array_equal(got, full[..., c0:c0 + nc]), rank
