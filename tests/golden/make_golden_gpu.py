"""Generates tests/golden/gpu_golden.npz from the UNMODIFIED reference's CUDA paths (oracle/_ref/libref.so:
Conv_gpu, backprop_gpu, backprop_gpu_cc, autoenc_fft, backprop_fft).  Needs a GPU, so it runs under gpurun:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/gpu_golden.npz'

and the result is copied to tests/golden/gpu_golden.npz and committed.  Square frames only for the training paths
(stride quirk C2: the compiled reference reads out of bounds when Nx != Ny)."""
import os
import sys

import re
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_lib as R  # noqa: E402


def capture_stdout(fn):
    """Run fn() with fd 1 redirected to a file; returns (result, text).  The reference prints its per-iteration
    "mse fft:" / "n: .. mse: .." values with cout (fft_backproplib.cu:1441,1464)."""
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tf:
        saved = os.dup(1)
        os.dup2(tf.fileno(), 1)
        try:
            res = fn()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tf.seek(0)
        return res, tf.read().decode()


def case(rng, dM, dD, Nk, Nl, Nx, Ny, wscale=0.2):
    inp = np.floor(rng.random((dD, Nx, Ny)) * 256).astype(np.float32)
    c = ((rng.random((dM, dD, Nk, Nl)) * 2 - 1) * wscale).astype(np.float32)
    f = ((rng.random((dD, dM, Nk, Nl)) * 2 - 1) * wscale).astype(np.float32)
    b = (rng.random(dM) * 2 - 1).astype(np.float32)
    p = (rng.random(dD) * 2 - 1).astype(np.float32)
    return inp, c, b, f, p


def main(out_path):
    g = {}
    rng = np.random.default_rng(11)
    # ---- Conv_gpu (any shape)
    for tag, dims in {"a": (4, 3, 5, 5, 20, 14), "b": (3, 2, 3, 3, 9, 13), "c": (2, 1, 7, 7, 16, 16)}.items():
        inp, c, b, _, _ = case(rng, *dims)
        g[f"convg_{tag}_x"], g[f"convg_{tag}_c"], g[f"convg_{tag}_b"] = inp, c, b
        g[f"convg_{tag}_out"] = R.conv_gpu(inp, c, b)
    # ---- backprop_gpu / backprop_gpu_cc, 2 consecutive steps on the same frame (momentum carried)
    for tag, dims in {"s5": (4, 3, 5, 5, 16, 16), "s3": (3, 2, 3, 3, 12, 12)}.items():
        for sym in (0, 1):
            inp, c, b, f, p = case(rng, *dims)
            if sym:
                f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
            hin = R.conv_gpu(inp, c, b)
            out = R.conv_gpu(hin, f, p)
            st = dict(c=c, b=b, f=f, p=p)
            for k in ("dc", "ddc"):
                st[k] = np.zeros_like(c)
            for k in ("df", "ddf"):
                st[k] = np.zeros_like(f)
            for k in ("db", "ddb"):
                st[k] = np.zeros_like(b)
            for k in ("dp", "ddp"):
                st[k] = np.zeros_like(p)
            key = f"bpg_{tag}_{sym}"
            for k, v in dict(inp=inp, hin=hin, out=out, **st).items():
                g[f"{key}_{k}"] = v
            for step in (1, 2):
                st = R.backprop_gpu(sym, inp, out, hin, st["c"], st["b"], st["f"], st["p"], st["dc"], st["db"], st["df"],
                                    st["dp"], st["ddc"], st["ddb"], st["ddf"], st["ddp"], 0.2, 0.9, 1)
                for k, v in st.items():
                    g[f"{key}_step{step}_{k}"] = v
    # ---- autoenc_fft: 1 pair and 2 pairs, fft_l = 1 (all layers) ; spectra cache returned
    for tag, (D, Nx, Ny, widths, sc) in {"p1": (3, 16, 16, [4], [2]), "p2": (2, 32, 32, [3, 5], [2, 2]),
                                         "s1": (3, 16, 16, [4], [1])}.items():
        x = np.floor(rng.random((D, Nx, Ny)) * 256).astype(np.float32)
        enc, dec, shapes_e, shapes_d = [], [], [], []
        d, nx, ny = D, Nx, Ny
        layer_shapes = [(D, Nx, Ny)]
        net_c, net_b, scale = [], [], []
        encs = []
        for w, s in zip(widths, sc):
            _, c, b, f, p = case(rng, w, d, 5, 5, 4, 4)
            encs.append((c, b, f, p, s, d, nx, ny))
            nx, ny = nx // s, ny // s
            layer_shapes += [(d, nx, ny), (w, nx, ny)]
            d = w
        for (c, b, f, p, s, d0, nx0, ny0) in reversed(encs):
            layer_shapes += [(d0, nx0 // s, ny0 // s), (d0, nx0, ny0)]
        net_c = [e[0] for e in encs] + [e[2] for e in reversed(encs)]
        net_b = [e[1] for e in encs] + [e[3] for e in reversed(encs)]
        scale = [e[4] for e in encs] + [-e[4] for e in reversed(encs)]
        layers, cfs = R.autoenc_fft(x, net_c, net_b, scale, layer_shapes, None, 1)
        # fft_l = 0: only the last layer is inverse-transformed.  (With fft_l = 1 cuFFT's multi-dimensional C2R
        # overwrites its INPUT spectrum, so every layer after the first inverse is computed from clobbered data.)
        layers0, _ = R.autoenc_fft(x, net_c, net_b, scale, layer_shapes, None, 0)
        g[f"aef_{tag}_last_fftl0"] = layers0[-1]
        layers0c, _ = R.autoenc_fft(x, net_c, net_b, scale, layer_shapes, cfs, 0)  # cached-spectra branch
        g[f"aef_{tag}_last_fftl0_cached"] = layers0c[-1]
        g[f"aef_{tag}_x"] = x
        g[f"aef_{tag}_scale"] = np.array(scale, np.int32)
        g[f"aef_{tag}_shapes"] = np.array(layer_shapes, np.int32)
        for n, (c, b) in enumerate(zip(net_c, net_b)):
            g[f"aef_{tag}_c{n}"], g[f"aef_{tag}_b{n}"] = c, b
            g[f"aef_{tag}_cf{n}"] = cfs[n]
        for l, a in enumerate(layers):
            g[f"aef_{tag}_L{l}"] = a
    # ---- backprop_fft: 100 iterations, with and without the multiobjective term
    # del0 = 0.2 is the app default (autoencoder.cpp:87); the small-del0 twins stay in the smooth regime, where 100
    # iterations of fp32 (reference) and fp64 (oracle) arithmetic do not drift apart.
    for tag, (dims, maxdiff, del0) in {"f5": ((4, 3, 5, 5, 16, 16), 0, 0.2), "f3": ((3, 2, 3, 3, 32, 16), 0, 0.2),
                                       "m5": ((4, 3, 5, 5, 16, 16), 1, 0.2), "g5": ((4, 3, 5, 5, 16, 16), 0, 0.005),
                                       "g3": ((3, 2, 3, 3, 32, 16), 0, 0.005), "n5": ((4, 3, 5, 5, 16, 16), 1, 0.005)}.items():
        inp, c, b, f, p = case(rng, *dims, wscale=0.5)
        dM, dD, Nk, Nl, Nx, Ny = dims
        shapes = [(dD, Nx, Ny), (dD, Nx, Ny), (dM, Nx, Ny), (dD, Nx, Ny), (dD, Nx, Ny)]
        layers, cfs = R.autoenc_fft(inp, [c, f], [b, p], [1, -1], shapes, None, 1)
        out = layers[3]
        res, text = capture_stdout(lambda: R.backprop_fft(layers[1], layers[1], out, cfs[0], c, cfs[1], f, b, p, del0, maxdiff))
        trace = [float(m) for m in re.findall(r"mse(?: fft)?: *([-+0-9.eEinfa]+)", text)]
        assert len(trace) == 101, (len(trace), text[:200])
        g[f"bpf_{tag}_trace"] = np.array(trace, np.float64)
        for k, v in dict(inp=layers[1], out=out, c=c, b=b, f=f, p=p, cfreq=cfs[0], ffreq=cfs[1]).items():
            g[f"bpf_{tag}_{k}"] = v
        for k, v in res.items():
            g[f"bpf_{tag}_new_{k}"] = v
        g[f"bpf_{tag}_maxdiff"] = np.int32(maxdiff)
        g[f"bpf_{tag}_del0"] = np.float32(del0)
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "gpu_golden.npz"))
