"""Reported baseline for the hand-written batched 2-D R2C / C2R (csrc/fft_kernels.cu): cuFFT on the same box at the batch
shapes of BASELINE config 3 (the reference's cufftPlanMany + cufftExecR2C / C2R call sites, fft_backproplib.cu:779-829),
through torch.fft (= cuFFT batched plans, plan cached after the first call), next to this engine's kernels through the C
ABI with device pointers.  Bytes counted = 4 P + 8 S per image (read once, write once).  Prints one JSON line.

    python tools/cufft_baseline.py            (on a B200; used by profiles/r2_fft_vs_cufft.md)
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "autoencoder-fft_b200"))


def main():
    import torch

    import aefft_ctypes as A

    dev = torch.device("cuda", 0)
    ctx = A.Ctx(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    A._chk(A.lib().aefft_set_stream(ctx.h, ctypes.c_void_p(stream.cuda_stream)))
    shapes = [(384, 1024, 1024), (2048, 512, 512), (4096, 256, 256), (8192, 128, 128), (8192, 64, 64)]
    rows = []
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts)

    for batch, Nx, Ny in shapes:
        x = torch.rand(batch, Nx, Ny, device=dev)
        spec = torch.empty(batch, Nx, Ny // 2 + 1, 2, device=dev)
        back = torch.empty_like(x)
        gb = batch * (4.0 * Nx * Ny + 8.0 * Nx * (Ny // 2 + 1)) / 1e9
        t_cufft_r2c = timeit(lambda: torch.fft.rfft2(x))
        X = torch.fft.rfft2(x)
        t_cufft_c2r = timeit(lambda: torch.fft.irfft2(X, s=(Nx, Ny), norm="forward"))
        t_r2c = timeit(lambda: ctx.fft_r2c(x, loc=A.DEVICE, out=spec))
        t_c2r = timeit(lambda: ctx.fft_c2r(spec, Ny, loc=A.DEVICE, out=back))
        err = float((torch.view_as_complex(spec) - X).abs().max() / X.abs().max())
        rows.append({"batch": batch, "Nx": Nx, "Ny": Ny, "GB": gb,
                     "cufft_r2c_ms": t_cufft_r2c, "cufft_c2r_ms": t_cufft_c2r, "aefft_r2c_ms": t_r2c, "aefft_c2r_ms": t_c2r,
                     "cufft_r2c_gbs": gb / t_cufft_r2c * 1e3, "cufft_c2r_gbs": gb / t_cufft_c2r * 1e3,
                     "aefft_r2c_gbs": gb / t_r2c * 1e3, "aefft_c2r_gbs": gb / t_c2r * 1e3, "max_rel_diff_vs_cufft": err})
    print(json.dumps({"what": "batched 2-D R2C / C2R, fp32, device-resident, best of 5 with an L2 flush between runs",
                      "rows": rows}))
    ctx.close()


if __name__ == "__main__":
    main()
