"""Reported baseline (not a target): the reference's own CUDA path -- Conv_gpu + backprop_gpu_cc of backproplib.cu,
rebuilt unmodified for sm_100 with nvcc 12.9 (oracle/_ref/libref.so) -- timed on this box through its public
nested-vector API, i.e. including its per-call packing, cudaMalloc/Free, H2D/D2H and dM*dD*Nk*Nl kernel launches with
thrust reductions (that IS its behaviour).  One frame per pair of BASELINE config 2, each pair at its own channel counts
on a SQUARE crop of its resolution (the reference reads out of bounds on non-square frames), scaled
linearly in the pixel count to the full frame.  Run by bench.py in a subprocess (a fault in the reference must not take
the bench down); prints one JSON line."""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_lib  # noqa: E402


def main():
    if not ref_lib.available():
        print(json.dumps({"unavailable": "oracle/_ref/libref.so not built"}))
        return
    pairs = [(3, 16, 320, 240), (16, 32, 160, 120), (32, 64, 80, 60)]
    rng = np.random.default_rng(0)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    total, parts = 0.0, []
    for dD, dM, Nx, Ny in pairs:
        n = min(Nx, Ny)
        x = np.floor(rng.random((dD, n, n)) * 256).astype(np.float32)
        c = ((rng.random((dM, dD, 5, 5)) * 2 - 1) * 0.3).astype(np.float32)
        f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
        b, p = np.zeros(dM, np.float32), np.zeros(dD, np.float32)
        z = [np.zeros_like(a) for a in (c, b, f, p, c, b, f, p)]
        os.dup2(devnull, 1)  # the reference prints "mse ..." lines
        try:
            best = None
            for rep in range(2):  # first repetition pays CUDA context / module load
                t0 = time.perf_counter()
                hin = ref_lib.conv_gpu(x, c, b)
                out = ref_lib.conv_gpu(hin, f, p)
                ref_lib.backprop_gpu(1, x, out, hin, c, b, f, p, *z, 0.2, 0.9)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        finally:
            os.dup2(saved, 1)
        scaled = best * (Nx * Ny) / (n * n)
        parts.append({"pair": f"{dD}->{dM}", "crop": f"{n}x{n}", "seconds": best, "scaled_seconds": scaled})
        total += scaled
    print(json.dumps({"value": 1.0 / total, "unit": "frames/s", "seconds_per_frame": total, "pairs": parts,
                      "what": "reference CUDA path (backproplib.cu rebuilt for sm_100, nvcc 12.9): Conv_gpu x2 + backprop_gpu_cc "
                              "per pair, 1 frame, square crops scaled linearly in pixels"}))


if __name__ == "__main__":
    main()
