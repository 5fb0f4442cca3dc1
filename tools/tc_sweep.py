"""Development probe: time the batched-over-bins tensor-core GEMM (aefft_spec_bin_gemm) at the c3 shapes, for ring depths."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "autoencoder-fft_b200"))
import numpy as np, torch
import aefft_ctypes as A

ctx = A.Ctx(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
A._chk(A.lib().aefft_set_stream(ctx.h, C.c_void_p(stream.cuda_stream)))
def run(S, ar, ac, amn, br, bc, bmn, M, N, K, outer, reps=5):
    a = torch.randn(S, ar, ac, device="cuda"); b = torch.randn(S, br, bc, device="cuda")
    out = torch.empty(S * M * N if not outer else S * M * N // 2, device="cuda")
    def go():
        A._chk(A.lib().aefft_spec_bin_gemm(ctx.h, C.c_int64(S), A._ptr(a), ar, ac, amn, A._ptr(b), br, bc, bmn, M, N, K, outer, 0, C.c_float(1.0), A._ptr(out)))
    go(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    gb = 4.0 * S * (ar * ac + br * bc + (M * N if not outer else M * N / 2)) / 1e9
    return min(ts), gb / min(ts) * 1e3
shapes = {"fwd p2 (128x128x64)": (8320, 128, 64, 0, 128, 64, 0, 128, 128, 64, 0),
          "fwd p1 (128x64x32)": (33024, 128, 32, 0, 64, 32, 0, 128, 64, 32, 0),
          "O   p2 (128x64x128)": (8320, 128, 128, 0, 64, 128, 0, 128, 64, 128, 0),
          "G   p2 (B mn)": (8320, 128, 64, 0, 64, 128, 1, 128, 128, 64, 0),
          "outer p2": (8320, 128, 128, 1, 128, 64, 1, 128, 64, 128, 1),
          "outer p1": (33024, 128, 64, 1, 128, 32, 1, 64, 32, 128, 1)}
for kn in os.environ.get("KNOCK", "0,1,2,4,7").split(","):
    os.environ["AEFFT_TC_KNOCK"] = kn
    for k, v in shapes.items():
        ms, gbs = run(*v)
        print(f"knock={kn} {k:24s} {ms:7.3f} ms {gbs:7.0f} GB/s (nominal bytes)")
