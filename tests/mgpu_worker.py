"""One rank of the multi-GPU parity tests (tests/test_multi_gpu.py): python tests/mgpu_worker.py MODE RANK WORLD IDFILE OUT.
Every rank owns GPU `RANK`, joins the engine's NCCL communicator (id passed through a file) and runs the same training
calls on its share of the frames; the trained weights / mse traces go to OUT (npz)."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "autoencoder-fft_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import aefft_ctypes as A  # noqa: E402
import oracle_np as O  # noqa: E402

CFG = dict(D=3, Nx=64, Ny=64, widths=[8, 16], pools=[2, 2], B_global=32, seed=99)


def main():
    mode, rank, world, idfile, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    ctx = A.Ctx(rank if world > 1 else 0)
    if world > 1:
        if rank == 0:
            ident = A.Ctx.comm_unique_id()
            with open(idfile + ".tmp", "wb") as fh:
                fh.write(ident)
            os.rename(idfile + ".tmp", idfile)
        else:
            for _ in range(600):
                if os.path.exists(idfile):
                    break
                time.sleep(0.1)
            ident = open(idfile, "rb").read()
        ctx.comm_init(ident, rank, world)
        if mode == "fft_bins":
            ctx.set_bin_shard(rank, world)
    B = CFG["B_global"] // world
    ctypes.CDLL("libc.so.6").srand(CFG["seed"])
    net = A.Net(ctx, CFG["D"], CFG["Nx"], CFG["Ny"], B)
    for m, s in zip(CFG["widths"], CFG["pools"]):
        net.add_layer(m, 1, 1, s, 0.1)
    x = O.synth_frames(5, B, CFG["D"], CFG["Nx"], CFG["Ny"], b0=rank * B)
    res = {}
    if mode == "coord":
        for n in range(net.num_pairs):
            net.set_symmetric(n)
        mse = np.zeros(64, np.float32)
        for _ in range(3):
            net.step(x, A.MODE_CUDA_REF_SYM, mse=mse)
        res["mse"] = mse[: net.num_pairs]
    else:
        tr = None
        for _ in range(2):
            tr = net.fft_step(x, del0=0.2, maxdiff=1 if mode == "fft_bins" else 0, n_iter=3, fft_l=-1)
        res["mse"] = tr
    for n in range(2 * net.num_pairs):
        c, b = net.get_conv(n)
        res[f"c{n}"], res[f"b{n}"] = c, b
    np.savez(out, **res)
    net.close()
    ctx.close()


if __name__ == "__main__":
    main()
