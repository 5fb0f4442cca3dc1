#!/usr/bin/env python
"""bench.py -- training frames/sec (forward + backprop + update) of the autoencoder hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c3]

Workloads (BASELINE.json configs; DESIGN.md "Measurement"):
  c2 (default, configs[1]): 3-pair coordinate-space autoencoder 3->16->32->64, 5x5 taps, pool 2 per pair, symmetric
      weights (backprop_gpu_cc semantics), 640x480 RGB frames, batch 64 per GPU.  One step = forward of the whole stack
      + one clipped-momentum update of every pair on the mean gradient of the batch.
  c3 (configs[2]): the same widths in momentum (FFT) space on 1024x1024 frames, batch 128 per GPU.
N>1: data-parallel frames (weak scaling: every rank owns its own batch), one NCCL all-reduce of the raw
kernel-space gradient block per pair, identical update on every rank.

Prints ONE JSON line (rank 0).  `value` = frames/s with frames resident in HBM; `e2e` = the same step through the
host-buffer C-ABI call (pinned host frames -> device every step, mse read back every step).
`--impl reference` times the reference's own CPU implementation (oracle/_ref/libref.so = unmodified netlib.cpp, else the
numpy port) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "autoencoder-fft_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

SEED = 1234
WORKLOADS = {
    # name: (D, Nx, Ny, widths, Lk, Ll, pool, rmax, batch per GPU, space)
    "c2": dict(D=3, Nx=640, Ny=480, widths=[16, 32, 64], Lk=1, Ll=1, pool=2, rmax=3.0, batch=64, space="coordinate"),
    "c3": dict(D=3, Nx=1024, Ny=1024, widths=[16, 32, 64], Lk=1, Ll=1, pool=2, rmax=3.0, batch=128, space="fft"),
    # configs[3]: 5 pairs (widths assumed), multiobjective term, 2048x2048 frames, FREQUENCY-BIN SHARDED:
    # every GPU holds all `batch` frames and owns a slab of spectrum columns (strong scaling); n_iter iterations per
    # backprop_fft call amortise the replicated forward / frame transforms (the reference runs 100 per call)
    "c4": dict(D=3, Nx=2048, Ny=2048, widths=[16, 32, 64, 128, 256], Lk=1, Ll=1, pool=2, rmax=3.0, batch=4, space="fft",
               shard="bins", n_iter=10, maxdiff=1),
}
DELMAX, ALPHA = 0.2, 0.9  # autoencoder.cpp:87-89


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


def pair_geometry(w):
    """[(dD, dM, Nx, Ny)] per pair (resolution at which the pair's convs run)."""
    out, d, nx, ny = [], w["D"], w["Nx"], w["Ny"]
    for m in w["widths"]:
        nx, ny = nx // w["pool"], ny // w["pool"]
        out.append((d, m, nx, ny))
        d = m
    return out


# ------------------------------------------------------------------------------------------------ reference CPU arm
def cpu_cost_units(dD, dM, T, P):
    """netlib.cpp:361-451 backprop: 9-deep loop nest, O(dM * dD^2 * (Nk*Nl)^2 * P)."""
    return float(dM) * dD * dD * T * T * P


def cpu_reference_sample(w, budget_s, repeats=1):
    """Times the reference's CPU path (Pool + Conv + Conv + backprop, netlib.cpp) on pair 0 of ONE frame, centre-cropped
    by Portion(q) so that one sample costs about budget_s, and extrapolates to the full step with the loop-nest cost
    model.  Returns dict(value frames/s, seconds per sample, sample description, kind, cores)."""
    import oracle_np as O
    import ref_lib

    kind = "reference" if ref_lib.available() else "port"
    geo = pair_geometry(w)
    Nk, Nl = 2 * (w["Lk"] + 1) + 1, 2 * (w["Ll"] + 1) + 1
    T = Nk * Nl
    dD, dM, nx, ny = geo[0]
    full_units = sum(cpu_cost_units(d, m, T, x * y) for d, m, x, y in geo)
    # measured on this image's Xeon: ~6.3e-6 s per cost unit (15.4 s for 8*1*625*307200 units)
    sec_per_unit = 6.3e-6 / 625.0
    q = 1
    while cpu_cost_units(dD, dM, T, (nx // q) * (ny // q)) * sec_per_unit > budget_s and min(nx, ny) // (2 * q) >= 16:
        q *= 2
    frame = O.synth_frames(SEED, 1, w["D"], w["Nx"], w["Ny"])[0]
    rng = O.GlibcRand(SEED)
    c, b = O.init_conv(rng, dM, dD, Nk, Nl, w["rmax"])
    f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
    p = np.zeros(dD, np.float32)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        if kind == "reference":
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)  # the reference prints "mse: ..." from inside backprop()
            try:
                pin = ref_lib.pool(frame, w["pool"], (nx, ny))
                pin, _, _ = ref_lib.portion(pin, pin, pin, q)
                hin = ref_lib.conv_cpu(pin, c, b)
                out = ref_lib.conv_cpu(hin, f, p)
                ref_lib.backprop_cpu(pin, out, hin, c, b, f, p, DELMAX)
            finally:
                os.dup2(saved, 1)
                os.close(devnull)
                os.close(saved)
        else:
            pin = O.pool(frame, w["pool"], (nx, ny))
            pin, _, _ = O.portion(pin, pin, pin, q)
            hin = O.conv_cpu(pin, c, b).astype(np.float32)
            out = O.conv_cpu(hin, f, p).astype(np.float32)
            O.backprop_cpu(pin, out, hin, c, b, f, p, DELMAX)
        times.append(time.perf_counter() - t0)
    t = min(times)
    sample_units = cpu_cost_units(dD, dM, T, (nx // q) * (ny // q))
    per_frame = t * full_units / sample_units
    desc = (f"1 frame, pair 0 only ({dD}->{dM}, {Nk}x{Nl}) on the centre {nx // q}x{ny // q} crop (Portion q={q}): Pool+Conv+Conv+"
            f"backprop of netlib.cpp, {t:.2f} s; extrapolated to all {len(geo)} pairs at full resolution with the loop-nest "
            f"cost dM*dD^2*(Nk*Nl)^2*P (x{full_units / sample_units:.0f})")
    return dict(value=1.0 / per_frame, seconds=t, sample=desc, kind=kind, cores=1, q=q)


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of
    this command (profiles/traffic.json, written from the .ncu-rep next to the summary it cites); None when no capture of
    that kernel is committed.  Compare with roofline.algorithmic_bytes_per_launch: traffic well above it = wasted re-reads."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tab = json.load(fh).get(workload, {})
    except (OSError, ValueError):
        return None
    key = "wgrad_ts" if kernel.startswith("wgrad_ts") else "conv_rs" if kernel.endswith("_rs") else kernel
    rec = tab.get(key)
    return float(rec["dram_bytes_per_launch"]) if rec else None


def ref_cuda_sample():
    """The reference's own CUDA kernels on this box (reported baseline): tools/ref_cuda_sample.py in a subprocess, so
    that a fault inside the reference cannot take the bench down.  None-like dict on failure."""
    import subprocess

    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_cuda_sample.py")], capture_output=True,
                             text=True, timeout=180)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"no result (rc={out.returncode}): {out.stderr.strip()[-200:]}"}
    except Exception as e:  # timeout, missing interpreter, ...
        return {"unavailable": repr(e)[:200]}


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    total_budget = 150.0
    per = max(0.5, total_budget / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu_reference_sample(w, per)
    res = [cpu_reference_sample(w, per) for _ in range(args.steps)]
    v = float(np.mean([r["value"] for r in res]))
    line = {
        "impl": "reference", "metric": "training frames/sec (fwd+backprop)", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean([r["seconds"] for r in res]) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, args, world),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": 1, "kind": res[0]["kind"], "sample": res[0]["sample"]},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def config_dict(w, args, world):
    geo = pair_geometry(w)
    return {
        "workload": f"{args.workload}: {len(w['widths'])}-pair {w['space']}-space autoencoder {w['D']}->" +
                    "->".join(map(str, w["widths"])) + f", {2 * (w['Lk'] + 1) + 1}x{2 * (w['Ll'] + 1) + 1} taps, pool {w['pool']}, "
                    f"symmetric weights, {w['Nx']}x{w['Ny']} frames, batch {args.batch} per GPU",
        "global_batch": args.batch * (1 if w.get("shard") == "bins" else world),
        "pairs": [{"dD": d, "dM": m, "Nx": x, "Ny": y} for d, m, x, y in geo],
        "parallelism": (f"bins{world} (frequency-bin sharded backprop_fft, {w.get('n_iter', 1)} iterations per call, "
                        f"forward replicated)") if w.get("shard") == "bins" else f"dp{world}",
        "cache": "inputs larger than L2 (frames + activations per step >> 126 MB); no explicit flush",
        "step": "forward of the full stack + gradients + clipped-momentum update of every pair",
    }


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CudaArray:
    """Wraps a raw device pointer so torch can view it (for torch.distributed.all_reduce on the gradient block)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class CoordWorkload:
    """c2: device-resident net, CUDA_REF_SYM step (forward + every pair), data-parallel all-reduce per pair."""

    def __init__(self, A, ctx, w, args, rank, world, dev, torch, dist):
        self.A, self.ctx, self.w, self.world, self.torch, self.dist, self.dev = A, ctx, w, world, torch, dist, dev
        B = self.B = args.batch
        ctypes.CDLL("libc.so.6").srand(SEED)
        self.net = net = A.Net(ctx, w["D"], w["Nx"], w["Ny"], B)
        for m in w["widths"]:
            net.add_layer(m, w["Lk"], w["Ll"], w["pool"], w["rmax"])
        self.P = net.num_pairs
        for n in range(self.P):
            net.set_symmetric(n)  # 'p' key: decoder = transposed encoder before symmetric training
        self.mode = A.MODE_CUDA_REF_SYM
        _, _, _, self.l0 = net.layer_info(0)
        self.n0 = B * w["D"] * w["Nx"] * w["Ny"]
        ctx.synth_frames(SEED, B, w["D"], w["Nx"], w["Ny"], b0=rank * B, out=self.l0, loc=A.DEVICE)
        self.gviews = []
        self.h2d_bytes, self.d2h_bytes = self.n0 * 4, 4 * self.P
        # end-to-end: pinned host frames, two device staging buffers, uploads on a second stream
        self.host = torch.empty(self.n0, dtype=torch.float32).pin_memory()
        A.lib().aefft_memcpy(ctx.h, ctypes.c_void_p(self.host.data_ptr()), ctypes.c_void_p(self.l0), ctypes.c_int64(self.n0 * 4), 1)
        self.mse_host = torch.zeros(64, dtype=torch.float32).pin_memory()
        self.stage = [torch.empty(self.n0, dtype=torch.float32, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        # byte-frame variant of the end-to-end path: the frames as the camera delivers them, interleaved 8-bit images
        # [B][rows = Ny][cols = Nx][D] (cv::Mat data); ImageToSpin_C runs on the device (aefft_net_set_frames_u8)
        f32 = self.host.numpy().reshape(B, w["D"], w["Nx"], w["Ny"])
        self.host_u8 = torch.from_numpy(np.ascontiguousarray(f32.transpose(0, 3, 2, 1)).astype(np.uint8)).pin_memory()
        self.stage_u8 = [torch.empty(self.host_u8.numel(), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.u8 = False

    def _train(self, frames_ptr, want_mse):
        A, net = self.A, self.net
        if self.world == 1:
            net.step(frames_ptr, self.mode, DELMAX, ALPHA, loc=A.DEVICE, mse=self.mse_host if want_mse else None)
            return
        net.forward(frames_ptr, loc=A.DEVICE)
        for n in range(self.P):
            ptr, glen = net.pair_gradients(n, self.mode)
            if len(self.gviews) <= n:
                self.gviews.append(self.torch.as_tensor(CudaArray(ptr, glen), device=self.dev))
            self.dist.all_reduce(self.gviews[n])
            net.pair_update(n, self.mode, self.B * self.world, DELMAX, ALPHA, want_mse=(want_mse and n == self.P - 1))

    def step_resident(self):
        self._train(None, False)

    def e2e_begin(self):
        """Queue the upload of the first batch."""
        self.k = 0
        self._upload(0)

    def _upload(self, slot):
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])  # the step that last read this buffer has finished
            if self.u8:
                self.stage_u8[slot].copy_(self.host_u8.view(-1), non_blocking=True)
            else:
                self.stage[slot].copy_(self.host, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def step_e2e(self):
        """One step on host frames: its own H2D (overlapped with the previous step's compute) + mse read back."""
        torch = self.torch
        slot = self.k % 2
        self._upload(1 - slot)  # next step's frames travel while this step computes
        torch.cuda.current_stream().wait_event(self.copied[slot])
        if self.u8:
            self.net.set_frames_u8(self.stage_u8[slot].data_ptr(), loc=self.A.DEVICE)
            self._train(None, True)
        else:
            self._train(self.stage[slot], True)
        self.consumed[slot].record(torch.cuda.current_stream())
        self.k += 1

    def describe_e2e(self):
        return ("pinned host frames -> double-buffered device staging on a copy stream (upload of step k+1 overlaps step k), "
                "aefft_net_step(AEFFT_DEVICE) + mse D2H and stream sync every step")

    def close(self):
        self.net.close()


class FftWorkload:
    """c3: momentum-space training.  One step = autoenc_fft forward of the whole stack (all layers materialised, as the
    reference needs them for training) + ONE iteration of backprop_fft's loop for every pair (the survey's
    definition of an FFT-space training step), on B frames, through the C ABI with device pointers."""

    def __init__(self, A, ctx, w, args, rank, world, dev, torch, dist):
        self.A, self.ctx, self.w, self.world, self.torch, self.dist = A, ctx, w, world, torch, dist
        B = self.B = args.batch

        Nk, Nl = 2 * (w["Lk"] + 1) + 1, 2 * (w["Ll"] + 1) + 1
        ctypes.CDLL("libc.so.6").srand(SEED)  # Init_conv draws from libc rand() (netlib.cpp:167-197), through the C ABI

        def init_conv(mS, dD_, kS, lS, rmax):
            c_ = np.empty((mS, dD_, kS, lS), np.float32)
            b_ = np.empty(mS, np.float32)
            A._chk(A.lib().aefft_init_conv(A._ptr(c_), A._ptr(b_), mS, dD_, kS, lS, ctypes.c_float(rmax)))
            return c_, b_

        encs, d, nx, ny = [], w["D"], w["Nx"], w["Ny"]
        shapes = [(d, nx, ny)]
        for m in w["widths"]:
            c, b = init_conv(m, d, Nk, Nl, w["rmax"] / 10.0)
            f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
            p_ = np.zeros(d, np.float32)
            encs.append((c, b, f, p_, d, nx, ny))
            nx, ny = nx // w["pool"], ny // w["pool"]
            shapes += [(d, nx, ny), (m, nx, ny)]
            d = m
        for (c, b, f, p_, d0, nx0, ny0) in reversed(encs):
            shapes += [(d0, nx0 // w["pool"], ny0 // w["pool"]), (d0, nx0, ny0)]
        net_c = [e[0] for e in encs] + [e[2] for e in reversed(encs)]
        net_b = [e[1] for e in encs] + [e[3] for e in reversed(encs)]
        self.scale = np.array([w["pool"]] * len(encs) + [-w["pool"]] * len(encs), np.int32)
        self.n_conv, self.shapes = len(net_c), shapes
        self.dims = np.array([x for c in net_c for x in c.shape], np.int32)
        self.coff = np.cumsum([0] + [c.size for c in net_c[:-1]]).astype(np.int64)
        self.boff = np.cumsum([0] + [b.size for b in net_b[:-1]]).astype(np.int64)
        self.ldims = np.array([x for s_ in shapes for x in s_], np.int32)
        lsz = [int(np.prod(s_)) for s_ in shapes]
        self.loff = np.cumsum([0] + lsz[:-1]).astype(np.int64)
        self.lstride = int(sum(lsz))
        self.c_all = ctx.to_device(np.concatenate([c.ravel() for c in net_c]))
        self.b_all = ctx.to_device(np.concatenate([b.ravel() for b in net_b]))
        self.layers = A.DevBuf(ctx, (B, self.lstride))
        self.n0 = lsz[0]
        frames = A.DevBuf(ctx, (B, self.n0))
        self.shard = w.get("shard") == "bins"
        self.n_iter, self.maxdiff = int(w.get("n_iter", 1)), int(w.get("maxdiff", 0))
        # data parallel: every rank owns its own frames; bin sharded: every rank holds the SAME frames
        ctx.synth_frames(SEED, B, w["D"], w["Nx"], w["Ny"], b0=0 if self.shard else rank * B, out=frames.ptr, loc=A.DEVICE)
        self.frames = frames
        self.host = torch.empty(B * self.n0, dtype=torch.float32).pin_memory()
        ctx.memcpy(self.host.data_ptr(), frames.ptr, B * self.n0 * 4, 1)
        self.trace = np.zeros(self.n_iter + 1, np.float32)
        self.pairs = []
        P = len(encs)
        for n in range(P):
            dM, dD = net_c[n].shape[:2]
            _, nx, ny = shapes[2 * n + 1]
            self.pairs.append(dict(dM=dM, dD=dD, Nx=nx, Ny=ny, Nk=Nk, Nl=Nl, l_in=2 * n + 1, l_out=len(shapes) - 2 - 2 * n,
                                   c=int(self.coff[n]), f=int(self.coff[2 * P - 1 - n]), b=int(self.boff[n]),
                                   p=int(self.boff[2 * P - 1 - n])))
        self.h2d_bytes, self.d2h_bytes = B * self.n0 * 4, 4 * P
        self._put_frames(A.DEVICE, frames.ptr)
        if world > 1 or self.shard:
            # data-parallel frames: every rank AVERAGES the raw kernel-space gradient block [dck | dfk | db | dp] of its own
            # frames over the ranks; bin sharded: every rank ADDS the partial block / mse of its spectrum columns (NCCL on
            # the engine's stream), in both cases before the clipped-momentum update
            views = {}
            op = dist.ReduceOp.SUM if self.shard else dist.ReduceOp.AVG

            def hook(ptr, n):
                if world == 1:
                    return
                if (ptr, n) not in views:
                    views[(ptr, n)] = torch.as_tensor(CudaArray(ptr, n), device=dev)
                dist.all_reduce(views[(ptr, n)], op=op)

            ctx.set_gradient_hook(hook)
            if self.shard:
                ctx.set_bin_shard(rank, world)

    def _put_frames(self, kind_loc, src_ptr):
        """layer 0 of every frame lives at layers[b*lstride]: strided copy of the batch's frames."""
        A, ctx = self.A, self.ctx
        A._chk(A.lib().aefft_memcpy2d(ctx.h, ctypes.c_void_p(self.layers.ptr), ctypes.c_int64(self.lstride * 4),
                                      ctypes.c_void_p(src_ptr), ctypes.c_int64(self.n0 * 4), ctypes.c_int64(self.n0 * 4),
                                      ctypes.c_int64(self.B), 0 if kind_loc == A.HOST else 2))

    def _train(self):
        A, ctx = self.A, self.ctx
        I32, I64, FP = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64), A.FP
        A._chk(A.lib().aefft_autoenc_fft(ctx.h, A.DEVICE, ctypes.c_int64(self.B), self.n_conv, self.dims.ctypes.data_as(I32),
                                         A._ptr(self.c_all), self.coff.ctypes.data_as(I64), A._ptr(self.b_all),
                                         self.boff.ctypes.data_as(I64), self.scale.ctypes.data_as(I32), len(self.shapes),
                                         self.ldims.ctypes.data_as(I32), A._ptr(self.layers), self.loff.ctypes.data_as(I64),
                                         ctypes.c_int64(self.lstride), 0, None, None, 1))
        raise_if = A._chk
        for q in self.pairs:
            # the pair's in/out layers are strided per frame inside `layers`; backprop_fft wants [B][dD][Nx][Ny] contiguous
            raise_if(A.lib().aefft_backprop_fft_strided(ctx.h, ctypes.c_int64(self.B), q["dD"], q["dM"], q["Nx"], q["Ny"], q["Nk"],
                                                        q["Nl"], ctypes.c_void_p(self.layers.ptr + int(self.loff[q["l_in"]]) * 4),
                                                        ctypes.c_void_p(self.layers.ptr + int(self.loff[q["l_out"]]) * 4),
                                                        ctypes.c_int64(self.lstride),
                                                        ctypes.c_void_p(self.c_all.ptr + q["c"] * 4), ctypes.c_void_p(self.c_all.ptr + q["f"] * 4),
                                                        ctypes.c_void_p(self.b_all.ptr + q["b"] * 4), ctypes.c_void_p(self.b_all.ptr + q["p"] * 4),
                                                        ctypes.c_float(DELMAX), self.maxdiff, self.n_iter, self.trace.ctypes.data_as(FP)))

    def step_resident(self):
        self._train()

    def e2e_begin(self):
        """Queue the upload of the first batch (double-buffered contiguous device staging on a copy stream)."""
        torch = self.torch
        if not hasattr(self, "stage"):
            self.stage = [torch.empty(self.B * self.n0, dtype=torch.float32, device=torch.cuda.current_device()) for _ in range(2)]
            self.copy_stream = torch.cuda.Stream()
            self.copied = [torch.cuda.Event() for _ in range(2)]
            self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.k = 0
        self._upload(0)

    def _upload(self, slot):
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            self.stage[slot].copy_(self.host, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def step_e2e(self):
        torch = self.torch
        slot = self.k % 2
        self._upload(1 - slot)  # next step's frames travel while this step computes
        torch.cuda.current_stream().wait_event(self.copied[slot])
        self._put_frames(self.A.DEVICE, self.stage[slot].data_ptr())  # into layer 0 of the per-frame blocks
        self.consumed[slot].record(torch.cuda.current_stream())
        self._train()  # ends with the mse trace D2H + stream sync inside aefft_backprop_fft
        self.k += 1

    def describe_e2e(self):
        return ("pinned host frames -> double-buffered device staging on a copy stream (upload of step k+1 overlaps step k) -> "
                "layer 0 of the per-frame blocks, mse trace read back per pair")

    def close(self):
        pass


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import aefft_ctypes as A

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = A.Ctx(local_rank)
    ctx.set_precision({"fp32": A.PRECISION_FP32, "bf16x3": A.PRECISION_BF16X3, "bf16": A.PRECISION_BF16}[args.precision])
    # all engine work, the NCCL collectives and the timing events share ONE non-default torch stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    A._chk(A.lib().aefft_set_stream(ctx.h, ctypes.c_void_p(stream.cuda_stream)))
    wl = (CoordWorkload if w["space"] == "coordinate" else FftWorkload)(A, ctx, w, args, rank, world, dev, torch, dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # started before the warm-up so that samples exist for short timed regions
    warm = max(args.warmup, 3)
    for _ in range(warm):
        wl.step_resident()
    barrier()
    A._chk(A.lib().aefft_profile_enable(ctx.h, 1))
    l_before = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wl.step_resident()
    e1.record()
    barrier()
    A._chk(A.lib().aefft_profile_enable(ctx.h, 0))
    launches = ctx.launches - l_before
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # per-kernel timing (CUDA events on the launching stream, inside the timed region)
    maxr = 64
    names = ctypes.create_string_buffer(64 * maxr)
    kms = (ctypes.c_float * maxr)()
    cnt = (ctypes.c_int64 * maxr)()
    fl = (ctypes.c_double * maxr)()
    by = (ctypes.c_double * maxr)()
    nrows = ctypes.c_int()
    A._chk(A.lib().aefft_profile_read(ctx.h, maxr, names, kms, cnt, fl, by, ctypes.byref(nrows)))
    kernels = []
    for k in range(nrows.value):
        nm = names.raw[64 * k: 64 * k + 64].split(b"\0")[0].decode()
        kernels.append(dict(name=nm, ms=float(kms[k]), launches=int(cnt[k]), flops=float(fl[k]), bytes=float(by[k])))
    kernels.sort(key=lambda r: -r["ms"])

    # ---- end to end: host frames every step, result read back every step
    wl.e2e_begin()
    for _ in range(2):
        wl.step_e2e()
    barrier()
    wl.e2e_begin()
    e0.record()
    for _ in range(args.steps):
        wl.step_e2e()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item())
    # ---- the same end-to-end step fed with 8-bit camera images (ImageToSpin_C on the device), coordinate space only
    e2e_u8_ms = None
    if hasattr(wl, "u8"):
        wl.u8 = True
        wl.e2e_begin()
        for _ in range(2):
            wl.step_e2e()
        barrier()
        wl.e2e_begin()
        e0.record()
        for _ in range(args.steps):
            wl.step_e2e()
        e1.record()
        barrier()
        ms3 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        e2e_u8_ms = float(ms3.item())
        wl.u8 = False
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        pk = peaks()
        B = args.batch
        sharded = w.get("shard") == "bins"  # every rank works on the same frames: total work is fixed (strong scaling)
        frames = B * (1 if sharded else world) * args.steps
        value = frames / (ms_total * 1e-3)
        top = kernels[0] if kernels else None
        roof = None
        if top and top["ms"] > 0:
            per_ms = top["ms"] / top["launches"]
            fl_l, by_l = top["flops"] / top["launches"], top["bytes"] / top["launches"]
            passes = 3 if (w["space"] == "coordinate" and args.precision == "bf16x3") else 1
            t_tensor = fl_l * passes / (pk["tf_sust"] * 1e12)  # fp32-grade results cost 3 bf16 MMA passes
            t_hbm = by_l / (pk["hbm"] * 1e9)
            if t_tensor >= t_hbm:
                ach = fl_l / (per_ms * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"]}
            else:
                ach = by_l / (per_ms * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
            if roof["bound"] == "tensor":
                roof.update({"mma_passes": passes, "pipe_frac": roof["frac"] * passes})
            roof.update({"traffic": ncu_traffic(args.workload, top["name"]), "kernel": top["name"], "avg_launch_ms": per_ms, "share_of_step": top["ms"] / ms_total,
                         "peak_source": pk["source"] + (", sustained bf16" if roof["bound"] == "tensor" else ""),
                         "algorithmic_flops_per_launch": fl_l, "algorithmic_bytes_per_launch": by_l})
        cfg = config_dict(w, args, world)
        cfg["e2e_path"] = wl.describe_e2e()
        line = {
            "metric": "training frames/sec (fwd+backprop)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "c64/f32" if w["space"] == "fft" else
                     {"fp32": "f32", "bf16x3": "f32 (bf16x3 split on tcgen05, fp32 accumulate)", "bf16": "bf16"}[args.precision],
            "data": "synthetic", "config": cfg,
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": wl.h2d_bytes,
                    "d2h_bytes_per_step": wl.d2h_bytes, "ms_per_step": e2e_ms / args.steps},
            "e2e_u8": None if e2e_u8_ms is None else {
                "value": frames / (e2e_u8_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": wl.h2d_bytes // 4,
                "d2h_bytes_per_step": wl.d2h_bytes, "ms_per_step": e2e_u8_ms / args.steps,
                "note": "same step, frames uploaded as interleaved 8-bit images (what the reference's camera delivers) and "
                        "converted on the device (aefft_net_set_frames_u8 = ImageToSpin_C); `e2e` above uploads fp32 frames"},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "roofline": roof,
            "kernels": [{"name": k["name"], "ms_per_step": k["ms"] / args.steps, "launches_per_step": k["launches"] / args.steps,
                         "tflops": (k["flops"] / (k["ms"] * 1e-3) / 1e12) if k["ms"] > 0 else None,
                         "gbs": (k["bytes"] / (k["ms"] * 1e-3) / 1e9) if k["ms"] > 0 else None} for k in kernels],
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_sample(w, 15.0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "frames/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"]}
            if w["space"] == "coordinate":
                line["ref_cuda_baseline"] = ref_cuda_sample()
        print(json.dumps(line))
    wl.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--size", type=int, default=None, help="square frames of this edge instead of the workload's (sweeps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.size:
        w["Nx"] = w["Ny"] = args.size
    if args.batch is None:
        args.batch = w["batch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
