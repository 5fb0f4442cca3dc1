#!/usr/bin/env python
"""bench.py -- training frames/sec (forward + backprop + update) of the autoencoder hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c1|c2|c3|c4] [--only]

Workloads (BASELINE.json configs; DESIGN.md "Measurement"):
  c1 (configs[0]): one pair 1->8, 5x5 taps, no pooling, 640x480 grayscale frame, CPU `backprop` semantics (CPU_REF mode:
      sequential-f update, no momentum) computed on the GPU.
  c2 (default headline, configs[1]): 3-pair coordinate-space autoencoder 3->16->32->64, 5x5 taps, pool 2 per pair,
      symmetric weights (backprop_gpu_cc semantics), 640x480 RGB frames, batch 64 per GPU.  One step = forward of the
      whole stack + one clipped-momentum update of every pair on the mean gradient of the batch.
  c3 (configs[2]): the same widths in momentum (FFT) space on 1024x1024 frames, batch 128 per GPU.
  c4 (configs[3]): 5 pairs, FFT space + multiobjective term, 2048x2048 frames, frequency-bin sharded over the GPUs.
The default run prints the c2 line and, inside it, `workloads.c3` (every N) and `workloads.c4` (N = 8): the FFT-space
half of the metric measured by the same command.  `--only` restricts the run to --workload.

N>1: one process per GPU (torch.distributed.run); the engine's own NCCL communicator (aefft_comm_init) all-reduces the
fused raw gradient block once per step; torch.distributed is used for rendezvous, barriers and max-over-ranks timing.

Prints ONE JSON line (rank 0).  `value` = frames/s with frames resident in HBM; `e2e` = the same step fed from pinned
HOST frames every step (H2D inside the timed region) with the mse read back every step.
`--impl reference` times the reference's own CPU implementation (oracle/_ref/libref.so = unmodified netlib.cpp, else the
numpy port) on a bounded sample of the same workload, on all host cores (one frame-sample per process; the reference
itself is single threaded).  Its `value` is what was MEASURED on the sample; the model-based projection to the full
configuration is a separate key (`extrapolated`) with its validation.
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

import dp  # noqa: E402  (who owns which frames / bins: the sharding plan the gloo tests check on CPU)

SEED = 1234
WORKLOADS = {
    "c1": dict(D=1, Nx=640, Ny=480, widths=[8], Lk=1, Ll=1, pool=1, rmax=3.0, batch=1, space="coordinate", mode="cpu_ref"),
    "c2": dict(D=3, Nx=640, Ny=480, widths=[16, 32, 64], Lk=1, Ll=1, pool=2, rmax=3.0, batch=64, space="coordinate"),
    "c3": dict(D=3, Nx=1024, Ny=1024, widths=[16, 32, 64], Lk=1, Ll=1, pool=2, rmax=3.0, batch=128, space="fft"),
    # configs[3]: 5 pairs (widths assumed), multiobjective term, 2048x2048 frames, FREQUENCY-BIN SHARDED:
    # every GPU holds all `batch` frames and owns a slab of spectrum columns (strong scaling); n_iter iterations per
    # backprop_fft call amortise the frame transforms (the reference runs 100 per call)
    "c4": dict(D=3, Nx=2048, Ny=2048, widths=[16, 32, 64, 128, 256], Lk=1, Ll=1, pool=2, rmax=3.0, batch=32, space="fft",
               shard="bins", n_iter=10, maxdiff=1),
    # the c2 stack in momentum space on the camera's 640x480 frames: lengths with factors 3 and 5 run on the
    # mixed-radix transforms (--workload c3cam --only; not part of the default line)
    "c3cam": dict(D=3, Nx=640, Ny=480, widths=[16, 32, 64], Lk=1, Ll=1, pool=2, rmax=3.0, batch=64, space="fft"),
}
DELMAX, ALPHA = 0.2, 0.9  # autoencoder.cpp:87-89
METRIC = "training frames/sec (fwd+backprop)"


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


def _silence_stdout():
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints "mse: ..." from inside backprop()
    return devnull, saved


def _restore_stdout(devnull, saved):
    os.dup2(saved, 1)
    os.close(devnull)
    os.close(saved)


def cpu_sample_once(D, dD, dM, Nx, Ny, pool, Lk, Ll, rmax, q, frame_index=0):
    """ONE execution of the reference's CPU path for one layer pair on one frame: Pool + Portion(q) + Conv + Conv +
    backprop (netlib.cpp), through oracle/_ref/libref.so when it was built (kind 'reference'), else the numpy port.
    Returns (seconds, kind, crop Nx, crop Ny)."""
    import oracle_np as O
    import ref_lib

    kind = "reference" if ref_lib.available() else "port"
    Nk, Nl = 2 * (Lk + 1) + 1, 2 * (Ll + 1) + 1
    frame = O.synth_frames(SEED, 1, dD, Nx * pool, Ny * pool, b0=frame_index)[0]
    rng = O.GlibcRand(SEED)
    c, b = O.init_conv(rng, dM, dD, Nk, Nl, rmax)
    f = np.ascontiguousarray(np.swapaxes(c, 0, 1))
    p = np.zeros(dD, np.float32)
    L = ref_lib if kind == "reference" else O
    t0 = time.perf_counter()
    devnull, saved = _silence_stdout()
    try:
        pin = L.pool(frame, pool, (Nx, Ny))
        pin, _, _ = L.portion(pin, pin, pin, q)
        hin = np.asarray(L.conv_cpu(pin, c, b), np.float32)
        out = np.asarray(L.conv_cpu(hin, f, p), np.float32)
        L.backprop_cpu(pin, out, hin, c, b, f, p, DELMAX)
    finally:
        _restore_stdout(devnull, saved)
    return time.perf_counter() - t0, kind, Nx // q, Ny // q


def _worker_main(argv):
    """`bench.py --cpu-worker D dD dM Nx Ny pool Lk Ll rmax q idx`: one sample in its own process (one host core)."""
    a = argv
    t, kind, cx, cy = cpu_sample_once(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6]), int(a[7]),
                                      float(a[8]), int(a[9]), int(a[10]))
    print(json.dumps({"seconds": t, "kind": kind, "crop": [cx, cy]}))


def cpu_parallel_samples(w, pair, q, n_proc):
    """n_proc independent frame-samples at once, one process per host core.  Returns (wall seconds, [per-process s], kind)."""
    dD, dM, nx, ny = pair
    procs = []
    t0 = time.perf_counter()
    for i in range(n_proc):
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(w["D"]), str(dD), str(dM), str(nx), str(ny),
               str(w["pool"]), str(w["Lk"]), str(w["Ll"]), str(w["rmax"]), str(q), str(i)]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    per, kind = [], "port"
    for pr in procs:
        out, _ = pr.communicate()
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith("{"):
                r = json.loads(ln)
                per.append(r["seconds"])
                kind = r["kind"]
                break
    return time.perf_counter() - t0, per, kind


SEC_PER_UNIT = 6.3e-6 / 625.0  # first guess only (15.4 s for 8*1*625*307200 units); re-fitted on every run


def choose_crop(w, pair, budget_s, sec_per_unit=SEC_PER_UNIT):
    dD, dM, nx, ny = pair
    T = (2 * (w["Lk"] + 1) + 1) * (2 * (w["Ll"] + 1) + 1)
    q = 1
    while cpu_cost_units(dD, dM, T, (nx // q) * (ny // q)) * sec_per_unit > budget_s and min(nx, ny) // (2 * q) >= 16:
        q *= 2
    return q


def cpu_reference_run(w, budget_s, cores=None, validate=True):
    """The reference's CPU path on a BOUNDED sample of workload `w`, on `cores` host cores (one frame-sample per process).
    value = frame-samples per second that were actually executed; `extrapolated` projects one full-configuration frame
    (all pairs, full resolution) with the loop-nest cost model, whose residual is measured on 2 crops x 2 channel counts."""
    geo = pair_geometry(w)
    Nk, Nl = 2 * (w["Lk"] + 1) + 1, 2 * (w["Ll"] + 1) + 1
    T = Nk * Nl
    cores = cores or max(1, min(os.cpu_count() or 1, 64))
    pair = geo[0]
    dD, dM, nx, ny = pair
    q = choose_crop(w, pair, budget_s)
    wall, per, kind = cpu_parallel_samples(w, pair, q, cores)
    if not per:
        return None
    t_mean = float(np.mean(per))
    sample_units = cpu_cost_units(dD, dM, T, (nx // q) * (ny // q))
    full_units = sum(cpu_cost_units(d, m, T, x * y) for d, m, x, y in geo)
    spu = t_mean / sample_units  # seconds per cost unit, fitted on this run's sample (loaded cores)
    res = dict(value=len(per) / wall, seconds=wall, per_process_seconds=per, kind=kind, cores=len(per), q=q,
               sample=(f"{len(per)} frame-samples in parallel (one process per host core, the reference is single threaded): "
                       f"each = pair 0 only ({dD}->{dM}, {Nk}x{Nl}) of one frame on the centre {nx // q}x{ny // q} crop "
                       f"(Portion q={q}): Pool+Conv+Conv+backprop of netlib.cpp; wall {wall:.2f} s, "
                       f"{t_mean:.2f} s per sample per core"),
               unit="frame-samples/s (see `sample`; NOT full-configuration frames)" if (q > 1 or len(geo) > 1) else "frames/s")
    extr = dict(value=len(per) / (spu * full_units * (wall / t_mean)), unit="frames/s (projected, full configuration)",
                model="seconds = k * sum_pairs dM*dD^2*(Nk*Nl)^2*Nx*Ny (netlib.cpp:361-451 loop nest), k fitted on the sample",
                factor=full_units / sample_units, seconds_per_frame_per_core=spu * full_units)
    if validate:
        # measured residual of the cost model across crop and channel changes (single process each, ~1 s per point)
        pts = []
        for (d2, m2) in ((dD, dM), (max(2, 2 * dD), max(4, dM // 2))):
            for q2 in (2 * q, 4 * q):
                if min(nx, ny) // q2 < 12:
                    continue
                t, _, cx, cy = cpu_sample_once(w["D"], d2, m2, nx, ny, w["pool"], w["Lk"], w["Ll"], w["rmax"], q2)
                pts.append((cpu_cost_units(d2, m2, T, cx * cy), t, f"{d2}->{m2}@{cx}x{cy}"))
        if len(pts) >= 2:
            k = float(np.exp(np.mean([np.log(t / u) for u, t, _ in pts])))
            resid = [float(t / (k * u) - 1.0) for u, t, _ in pts]
            extr["validation"] = {"points": [p_[2] for p_ in pts], "seconds": [p_[1] for p_ in pts],
                                  "relative_residuals": resid, "max_abs_residual": float(np.max(np.abs(resid))),
                                  "k_single_process": k, "k_sample": spu}
    res["extrapolated"] = extr
    return res


def fft_port_sample(w, budget_s):
    """FFT-space workloads: the reference's momentum path needs cuFFT + a GPU (it has no CPU mode), so the CPU arm is the
    numpy PORT (oracle_np) of autoenc_fft + one backprop_fft iteration per pair on ONE frame at reduced resolution."""
    import oracle_np as O

    Nk, Nl = 2 * (w["Lk"] + 1) + 1, 2 * (w["Ll"] + 1) + 1
    size = 128
    rng = O.GlibcRand(SEED)
    encs, d = [], w["D"]
    for m in w["widths"][:3]:
        c, b = O.init_conv(rng, m, d, Nk, Nl, w["rmax"] / 10.0)
        encs.append((c, b, np.ascontiguousarray(np.swapaxes(c, 0, 1)), np.zeros(d, np.float32)))
        d = m
    net_c = [e[0] for e in encs] + [e[2] for e in reversed(encs)]
    net_b = [e[1] for e in encs] + [e[3] for e in reversed(encs)]
    scale = [w["pool"]] * len(encs) + [-w["pool"]] * len(encs)
    x = O.synth_frames(SEED, 1, w["D"], size, size)[0]
    t0 = time.perf_counter()
    layers, _ = O.autoenc_fft(x, net_c, net_b, scale, None, 1)
    P = len(encs)
    for n in range(P):
        c, b, f, p = encs[n]
        O.backprop_fft(layers[2 * n + 1], layers[2 * n + 1], layers[len(layers) - 2 - 2 * n], c, f, b, p, DELMAX,
                       int(w.get("maxdiff", 0)) if c.shape[0] * c.shape[1] <= 512 else 0, 1)
    t = time.perf_counter() - t0
    return dict(value=1.0 / t, seconds=t, kind="port", cores=1, unit="frame-samples/s (see `sample`)",
                sample=f"numpy port (oracle_np) of autoenc_fft + 1 backprop_fft iteration per pair, first {P} pairs, ONE frame at "
                       f"{size}x{size} instead of {w['Nx']}x{w['Ny']}: {t:.2f} s (the reference's FFT path has no CPU mode)")


def cpu_baseline_for(w, budget_s, validate=True):
    return fft_port_sample(w, budget_s) if w["space"] == "fft" else cpu_reference_run(w, budget_s, validate=validate)


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


def ref_cuda_sample(extra=()):
    """The reference's own CUDA kernels on this box (reported baseline): tools/ref_cuda_sample.py in a subprocess, so
    that a fault inside the reference cannot take the bench down.  None-like dict on failure."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_cuda_sample.py"), *extra], capture_output=True,
                             text=True, timeout=240)
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
    n_runs = max(1, args.steps + args.warmup)
    per = max(0.5, min(20.0, total_budget / n_runs))
    for _ in range(args.warmup):
        cpu_baseline_for(w, per, validate=False)
    res = [cpu_baseline_for(w, per, validate=(i == 0)) for i in range(args.steps)]
    res = [r for r in res if r]
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "the reference CPU sample did not run"}))
        return
    v = float(np.mean([r["value"] for r in res]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s",
        "value_is": res[0]["unit"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean([r["seconds"] for r in res]) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, args, world),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": res[0]["cores"], "kind": res[0]["kind"],
                         "sample": res[0]["sample"]},
        "extrapolated": res[0].get("extrapolated"),
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def config_dict(w, args, world, name=None, batch=None):
    geo = pair_geometry(w)
    batch = batch if batch is not None else w["batch"]
    return {
        "workload": f"{name or args.workload}: {len(w['widths'])}-pair {w['space']}-space autoencoder {w['D']}->" +
                    "->".join(map(str, w["widths"])) + f", {2 * (w['Lk'] + 1) + 1}x{2 * (w['Ll'] + 1) + 1} taps, pool {w['pool']}, "
                    + ("CPU backprop semantics (sequential-f, no momentum), " if w.get("mode") == "cpu_ref" else "symmetric weights, ")
                    + f"{w['Nx']}x{w['Ny']} frames, batch {batch} per GPU",
        "global_batch": batch * (1 if w.get("shard") == "bins" else world),
        "pairs": [{"dD": d, "dM": m, "Nx": x, "Ny": y} for d, m, x, y in geo],
        "parallelism": (f"bins{world} (forward data parallel over {world} x {batch // world if not name or True else batch} frames, all-to-all of the "
                        f"pairs' spectrum slabs, frequency-bin sharded backprop_fft with {w.get('n_iter', 1)} iterations per call)")
                       if w.get("shard") == "bins" else f"dp{world}",
        "cache": "inputs larger than L2 (frames + activations per step >> 126 MB); no explicit flush" if batch * w["Nx"] * w["Ny"] * w["D"] * 4 > 2e8
                 else "L2 flushed between timed iterations (256 MB scratch write)",
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


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this process (and so its pinned frame buffers, first touched after this call) to the NUMA node the GPU hangs
    off: at N>1 the H2D copies of all ranks otherwise leave from node 0's memory controllers.  Best effort."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA information for the GPU"}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = ids & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": cpus, "bound": bool(allowed)}
    except Exception as e:  # no sysfs, no permission, ...
        return {"numa_node": None, "note": repr(e)[:120]}


class Staging:
    """End-to-end input path shared by the workloads: pinned host frames -> double-buffered device staging on a copy
    stream (the upload of step k+1 overlaps step k)."""

    def __init__(self, torch, dev, host_tensor):
        self.torch, self.host = torch, host_tensor
        self.stage = [torch.empty(host_tensor.numel(), dtype=host_tensor.dtype, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.k = 0

    def begin(self):
        self.k = 0
        self.upload(0)

    def upload(self, slot):
        torch = self.torch
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])  # the step that last read this buffer has finished
            self.stage[slot].copy_(self.host.view(-1), non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def acquire(self):
        """Device buffer holding this step's frames (its upload was queued one step ago); queues the next upload."""
        slot = self.k % 2
        self.upload(1 - slot)
        self.torch.cuda.current_stream().wait_event(self.copied[slot])
        return slot, self.stage[slot]

    def release(self, slot):
        self.consumed[slot].record(self.torch.cuda.current_stream())
        self.k += 1


class CoordWorkload:
    """c2 / c1: device-resident net; c2 = CUDA_REF_SYM step (forward + every pair), c1 = CPU_REF semantics.  Data
    parallel: the engine all-reduces the fused gradient block of all pairs once per step (aefft_comm_init)."""

    def __init__(self, A, ctx, w, batch, rank, world, dev, torch):
        self.A, self.ctx, self.w, self.world, self.torch, self.dev = A, ctx, w, world, torch, dev
        B = self.B = batch
        ctypes.CDLL("libc.so.6").srand(SEED)
        self.net = net = A.Net(ctx, w["D"], w["Nx"], w["Ny"], B)
        for m in w["widths"]:
            net.add_layer(m, w["Lk"], w["Ll"], w["pool"], w["rmax"])
        self.P = net.num_pairs
        self.mode = A.MODE_CPU_REF if w.get("mode") == "cpu_ref" else A.MODE_CUDA_REF_SYM
        if self.mode == A.MODE_CUDA_REF_SYM:
            for n in range(self.P):
                net.set_symmetric(n)  # 'p' key: decoder = transposed encoder before symmetric training
        _, _, _, self.l0 = net.layer_info(0)
        self.n0 = B * w["D"] * w["Nx"] * w["Ny"]
        ctx.synth_frames(SEED, B, w["D"], w["Nx"], w["Ny"], b0=dp.frame_range(rank, world, B)[0], out=self.l0, loc=A.DEVICE)
        self.h2d_bytes, self.d2h_bytes = self.n0 * 4, 4 * self.P
        host = torch.empty(self.n0, dtype=torch.float32).pin_memory()
        A.lib().aefft_memcpy(ctx.h, ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(self.l0), ctypes.c_int64(self.n0 * 4), 1)
        self.mse_host = torch.zeros(64, dtype=torch.float32).pin_memory()
        self.f32 = Staging(torch, dev, host)
        # byte-frame variant of the end-to-end path: the frames as the camera delivers them, interleaved 8-bit images
        # [B][rows = Ny][cols = Nx][D] (cv::Mat data); ImageToSpin_C runs on the device (aefft_net_set_frames_u8)
        f32 = host.numpy().reshape(B, w["D"], w["Nx"], w["Ny"])
        host_u8 = torch.from_numpy(np.ascontiguousarray(f32.transpose(0, 3, 2, 1)).astype(np.uint8)).pin_memory()
        self.u8s = Staging(torch, dev, host_u8)
        self.u8 = False
        self.has_u8 = w["D"] <= 4

    def _train(self, frames_ptr, want_mse):
        self.net.step(frames_ptr, self.mode, DELMAX, ALPHA, loc=self.A.DEVICE, mse=self.mse_host if want_mse else None)

    def step_resident(self):
        self._train(None, False)

    def e2e_begin(self):
        (self.u8s if self.u8 else self.f32).begin()

    def step_e2e(self):
        """One step on host frames: its own H2D (overlapped with the previous step's compute) + mse read back."""
        st = self.u8s if self.u8 else self.f32
        slot, buf = st.acquire()
        if self.u8:
            self.net.set_frames_u8(buf.data_ptr(), loc=self.A.DEVICE)
            self._train(None, True)
        else:
            self._train(buf, True)
        st.release(slot)

    def describe_e2e(self):
        return ("pinned host frames -> double-buffered device staging on a copy stream (upload of step k+1 overlaps step k), "
                "aefft_net_step(AEFFT_DEVICE) + mse D2H and stream sync every step")

    def close(self):
        self.net.close()


class FftNetWorkload:
    """c3: momentum-space training on the device-resident net (aefft_net_fft_step): one R2C of the frames, forward of the
    whole stack with every layer's spectrum kept in HBM, ONE iteration of backprop_fft's loop for every pair directly on
    those spectra (the survey's definition of an FFT-space training step), one C2R of the reconstruction (fft_l = 0)."""

    def __init__(self, A, ctx, w, batch, rank, world, dev, torch):
        self.A, self.ctx, self.w, self.world, self.torch = A, ctx, w, world, torch
        self.shard = w.get("shard") == "bins"
        if self.shard:
            # c4: the GLOBAL batch is fixed (strong scaling): every rank forwards batch/world frames, the pairs' spectra are
            # exchanged into column slabs (all-to-all) and training runs bin sharded (aefft_set_bin_shard)
            assert batch % world == 0, "c4: the batch must divide over the ranks"
            batch //= world
            if world > 1:
                ctx.set_bin_shard(rank, world)
        B = self.B = batch
        ctypes.CDLL("libc.so.6").srand(SEED)
        self.net = net = A.Net(ctx, w["D"], w["Nx"], w["Ny"], B)
        for m in w["widths"]:
            net.add_layer(m, w["Lk"], w["Ll"], w["pool"], w["rmax"] / 10.0)
        self.P = net.num_pairs
        for n in range(self.P):
            net.set_symmetric(n)
        self.n_iter, self.maxdiff = int(w.get("n_iter", 1)), int(w.get("maxdiff", 0))
        _, _, _, self.l0 = net.layer_info(0)
        self.n0 = B * w["D"] * w["Nx"] * w["Ny"]
        ctx.synth_frames(SEED, B, w["D"], w["Nx"], w["Ny"], b0=dp.frame_range(rank, world, B)[0], out=self.l0, loc=A.DEVICE)
        host = torch.empty(self.n0, dtype=torch.float32).pin_memory()
        ctx.memcpy(host.data_ptr(), self.l0, self.n0 * 4, 1)
        self.f32 = Staging(torch, dev, host)
        self.h2d_bytes, self.d2h_bytes = self.n0 * 4, 4 * self.P * (self.n_iter + 1)
        # 8-bit camera frames [B][rows = Ny][cols = Nx][D] (cv::Mat data), converted on the device (ImageToSpin_C)
        f32 = host.numpy().reshape(B, w["D"], w["Nx"], w["Ny"])
        host_u8 = torch.from_numpy(np.ascontiguousarray(f32.transpose(0, 3, 2, 1)).astype(np.uint8)).pin_memory()
        self.u8s = Staging(torch, dev, host_u8)
        self.u8 = False
        self.has_u8 = w["D"] <= 4

    def step_resident(self):
        self.net.fft_step(None, DELMAX, self.maxdiff, self.n_iter, fft_l=0, loc=self.A.DEVICE, want_mse=False)

    def e2e_begin(self):
        (self.u8s if self.u8 else self.f32).begin()

    def step_e2e(self):
        st = self.u8s if self.u8 else self.f32
        slot, buf = st.acquire()
        if self.u8:
            self.net.set_frames_u8(buf.data_ptr(), loc=self.A.DEVICE)
            buf = None
        self.net.fft_step(buf, DELMAX, self.maxdiff, self.n_iter, fft_l=0, loc=self.A.DEVICE, want_mse=True)  # syncs (mse D2H)
        st.release(slot)

    def describe_e2e(self):
        return ("pinned host frames -> double-buffered device staging on a copy stream (upload of step k+1 overlaps step k) -> "
                "aefft_net_fft_step(AEFFT_DEVICE), mse traces of all pairs read back every step")

    def close(self):
        self.net.close()


class FftWorkload:
    """c3 (reference-shaped C-ABI path) / c4: momentum-space training.  One step = autoenc_fft forward of the whole stack (all layers materialised, as
    the reference needs them for training) + n_iter iterations of backprop_fft's loop for every pair (c3: ONE,
    the survey's definition of an FFT-space training step), on B frames, through the C ABI with device pointers."""

    def __init__(self, A, ctx, w, batch, rank, world, dev, torch):
        self.A, self.ctx, self.w, self.world, self.torch = A, ctx, w, world, torch
        B = self.B = batch
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
        ctx.synth_frames(SEED, B, w["D"], w["Nx"], w["Ny"], b0=0 if self.shard else dp.frame_range(rank, world, B)[0], out=frames.ptr, loc=A.DEVICE)
        self.frames = frames
        host = torch.empty(B * self.n0, dtype=torch.float32).pin_memory()
        ctx.memcpy(host.data_ptr(), frames.ptr, B * self.n0 * 4, 1)
        self.f32 = Staging(torch, dev, host)
        self.trace = np.zeros(self.n_iter + 1, np.float32)
        self.pairs = []
        P = len(encs)
        for n in range(P):
            dM, dD = net_c[n].shape[:2]
            _, nx, ny = shapes[2 * n + 1]
            self.pairs.append(dict(dM=dM, dD=dD, Nx=nx, Ny=ny, Nk=Nk, Nl=Nl, l_in=2 * n + 1, l_out=len(shapes) - 2 - 2 * n,
                                   c=int(self.coff[n]), f=int(self.coff[2 * P - 1 - n]), b=int(self.boff[n]),
                                   p=int(self.boff[2 * P - 1 - n])))
        self.h2d_bytes, self.d2h_bytes = B * self.n0 * 4, 4 * P * (self.n_iter + 1)
        self._put_frames(A.DEVICE, frames.ptr)
        # data-parallel frames: the engine AVERAGES the raw kernel-space gradient block over the ranks (its own NCCL
        # all-reduce on the ctx stream); bin sharded: it ADDS the partial blocks of the ranks' spectrum columns
        if self.shard and world > 1:
            ctx.set_bin_shard(rank, world)
        self.has_u8 = False

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
        for q in self.pairs:
            # the pair's in/out layers are strided per frame inside `layers`
            A._chk(A.lib().aefft_backprop_fft_strided(ctx.h, ctypes.c_int64(self.B), q["dD"], q["dM"], q["Nx"], q["Ny"], q["Nk"],
                                                      q["Nl"], ctypes.c_void_p(self.layers.ptr + int(self.loff[q["l_in"]]) * 4),
                                                      ctypes.c_void_p(self.layers.ptr + int(self.loff[q["l_out"]]) * 4),
                                                      ctypes.c_int64(self.lstride),
                                                      ctypes.c_void_p(self.c_all.ptr + q["c"] * 4), ctypes.c_void_p(self.c_all.ptr + q["f"] * 4),
                                                      ctypes.c_void_p(self.b_all.ptr + q["b"] * 4), ctypes.c_void_p(self.b_all.ptr + q["p"] * 4),
                                                      ctypes.c_float(DELMAX), self.maxdiff, self.n_iter, self.trace.ctypes.data_as(FP)))

    def step_resident(self):
        self._train()

    def e2e_begin(self):
        self.f32.begin()

    def step_e2e(self):
        slot, buf = self.f32.acquire()
        self._put_frames(self.A.DEVICE, buf.data_ptr())  # into layer 0 of the per-frame blocks
        self.f32.release(slot)
        self._train()  # ends with the mse trace D2H + stream sync inside aefft_backprop_fft

    def describe_e2e(self):
        return ("pinned host frames -> double-buffered device staging on a copy stream (upload of step k+1 overlaps step k) -> "
                "layer 0 of the per-frame blocks, mse trace read back per pair")

    def close(self):
        self.layers.free()
        self.frames.free()
        self.c_all.free()
        self.b_all.free()


def roofline_of(top, w, precision, pk, ms_total, workload):
    if not top or top["ms"] <= 0:
        return None
    per_ms = top["ms"] / top["launches"]
    fl_l, by_l = top["flops"] / top["launches"], top["bytes"] / top["launches"]
    tc_kernel = top["name"].endswith(("_rs", "_ts", "_tc"))
    passes = 3 if (tc_kernel and precision == "bf16x3") else 1
    t_tensor = fl_l * passes / (pk["tf_sust"] * 1e12) if tc_kernel else 0.0  # fp32-grade results cost 3 MMA passes
    t_hbm = by_l / (pk["hbm"] * 1e9)
    if t_tensor >= t_hbm:
        ach = fl_l / (per_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                "mma_passes": passes, "pipe_frac": ach / pk["tf_sust"] * passes}
    else:
        ach = by_l / (per_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    roof.update({"traffic": ncu_traffic(workload, top["name"]), "kernel": top["name"], "avg_launch_ms": per_ms,
                 "share_of_step": top["ms"] / ms_total,
                 "peak_source": pk["source"] + (", sustained bf16" if roof["bound"] == "tensor" else ""),
                 "algorithmic_flops_per_launch": fl_l, "algorithmic_bytes_per_launch": by_l})
    return roof


def measure(env, name, w, batch, steps, warmup, precision):
    """One workload on this rank's GPU: resident-frames timing (per-kernel events inside), end-to-end timing from host
    frames.  Returns the record (rank 0) or None."""
    torch, dist, A = env["torch"], env["dist"], env["A"]
    rank, world, dev, local_rank = env["rank"], env["world"], env["dev"], env["local_rank"]
    ctx = A.Ctx(local_rank)
    ctx.set_precision({"fp32": A.PRECISION_FP32, "bf16x3": A.PRECISION_BF16X3, "bf16": A.PRECISION_BF16}[precision])
    stream = env["stream"]
    A._chk(A.lib().aefft_set_stream(ctx.h, ctypes.c_void_p(stream.cuda_stream)))
    if world > 1:
        # the engine's own communicator: rank 0's NCCL unique id travels through torch.distributed (rendezvous only).
        # NCCL prints its version banner to stdout on the first communicator: keep stdout for the one JSON line.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            ident = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                ident.copy_(torch.frombuffer(bytearray(A.Ctx.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(ident, 0)
            ctx.comm_init(bytes(ident.cpu().numpy().tobytes()), rank, world)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    cls = CoordWorkload if w["space"] == "coordinate" else FftWorkload if env.get("fft_capi") else FftNetWorkload
    wl = cls(A, ctx, w, batch, rank, world, dev, torch)
    small = batch * w["Nx"] * w["Ny"] * w["D"] * 4 <= 2e8  # working set may sit in the 126 MB L2: flush between iterations
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev) if small else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n steps bracketed by barrier + synchronize; device time of the steps only (the L2 flush, when used, sits
        between event pairs)."""
        barrier()
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
        else:
            evs = []
            for _ in range(n):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warm = max(warmup, 3)
    for _ in range(warm):
        wl.step_resident()
    barrier()
    # Two timed regions of K steps each: the first with nothing but the steps on the stream (`value`), the second with the
    # library's per-launch CUDA events switched on (two event records around every kernel: the per-kernel table and the
    # roofline's launch duration come from this one; its step time is reported as ms_per_step_instrumented).
    l_before = ctx.launches
    ms_total = timed(wl.step_resident, steps)
    launches = ctx.launches - l_before
    ctx.profile_enable(True)
    ms_prof = timed(wl.step_resident, steps)
    ctx.profile_enable(False)
    kernels = ctx.profile_records(96)
    kernels.sort(key=lambda r: -r["ms"])

    def e2e_run():
        wl.e2e_begin()
        for _ in range(2):
            wl.step_e2e()
        barrier()
        wl.e2e_begin()
        return timed(wl.step_e2e, steps)

    e2e_ms = e2e_run()
    e2e_u8_ms = None
    if getattr(wl, "has_u8", False):
        wl.u8 = True
        e2e_u8_ms = e2e_run()
        wl.u8 = False
    rec = None
    if rank == 0:
        pk = peaks()
        sharded = w.get("shard") == "bins"  # every rank works on the same frames: total work is fixed (strong scaling)
        frames = batch * (1 if sharded else world) * steps
        cfg = config_dict(w, None, world, name=name, batch=batch)
        cfg["e2e_path"] = wl.describe_e2e()
        rec = {
            "metric": METRIC, "value": frames / (ms_total * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms_total / steps,
            "ms_per_step_instrumented": ms_prof / steps,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "c64/f32" if w["space"] == "fft" else
                     {"fp32": "f32", "bf16x3": "f32 (bf16x3 split on tcgen05, fp32 accumulate)", "bf16": "bf16"}[precision],
            "data": "synthetic", "config": cfg,
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": wl.h2d_bytes,
                    "d2h_bytes_per_step": wl.d2h_bytes, "ms_per_step": e2e_ms / steps,
                    "h2d_gbs_per_gpu": wl.h2d_bytes / (e2e_ms / steps * 1e-3) / 1e9},
            "e2e_u8": None if e2e_u8_ms is None else {
                "value": frames / (e2e_u8_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": wl.h2d_bytes // 4,
                "d2h_bytes_per_step": wl.d2h_bytes, "ms_per_step": e2e_u8_ms / steps,
                "note": "same step, frames uploaded as interleaved 8-bit images (what the reference's camera delivers) and "
                        "converted on the device (aefft_net_set_frames_u8 = ImageToSpin_C); `e2e` above uploads fp32 frames"},
            "path": cls.__name__ + ": " + " ".join((cls.__doc__ or "").split())[:400],
            "gpu_launches": int(launches) * world,
            "collectives_per_step": (0 if world == 1 else 1 if w["space"] == "coordinate" else
                                     len(w["widths"]) * (int(w.get("n_iter", 1)) + 1)),
            "roofline": roofline_of(kernels[0] if kernels else None, w, precision, pk, ms_prof, name),
            "kernels": [{"name": k["name"], "ms_per_step": k["ms"] / steps, "launches_per_step": k["launches"] / steps,
                         "tflops": (k["flops"] / (k["ms"] * 1e-3) / 1e12) if k["ms"] > 0 else None,
                         "gbs": (k["bytes"] / (k["ms"] * 1e-3) / 1e9) if k["ms"] > 0 else None} for k in kernels],
        }
    wl.close()
    del wl, flush
    ctx.close()
    torch.cuda.empty_cache()
    return rec


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import aefft_ctypes as A

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else {"numa_node": None, "note": "single process: not bound"}
    if world > 1:
        # NCCL prints its version banner to stdout when the first communicator comes up: stdout carries only the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    # all engine work, the collectives and the timing events share ONE non-default torch stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    env = dict(torch=torch, dist=dist, A=A, rank=rank, world=world, dev=dev, local_rank=local_rank, stream=stream,
               fft_capi=args.fft_capi)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # started before the warm-up so that samples exist for short timed regions
    w = dict(WORKLOADS[args.workload])
    if args.size:
        w["Nx"] = w["Ny"] = args.size
    batch = args.batch if args.batch is not None else w["batch"]
    line = measure(env, args.workload, w, batch, args.steps, args.warmup, args.precision)
    extra = {}
    if not args.only and args.workload == "c2" and not args.size:
        # the FFT-space half of the metric, measured by the same command: c3 at every N, c4 (bin sharded) at N = 8
        for nm in (["c3"] + (["c4"] if world == 8 else [])):
            try:
                ww = dict(WORKLOADS[nm])
                extra[nm] = measure(env, nm, ww, ww["batch"], args.steps if nm == "c3" else max(2, min(args.steps, 5)),
                                    args.warmup, args.precision)
            except Exception as e:  # a failing side workload must not take the headline down
                extra[nm] = {"error": repr(e)[:300]}
    if not args.only and args.workload == "c2" and not args.size and not args.fft_capi:
        # the same c3 step through the reference-shaped C-ABI calls (aefft_autoenc_fft with fft_l = 1, then
        # aefft_backprop_fft on the real-space layers, i.e. with the per-layer C2R / R2C round trips the reference makes)
        try:
            env2 = dict(env, fft_capi=True)
            ww = dict(WORKLOADS["c3"])
            r = measure(env2, "c3", ww, ww["batch"], max(2, min(args.steps, 5)), args.warmup, args.precision)
            if rank == 0 and isinstance(extra.get("c3"), dict) and "error" not in extra["c3"]:
                extra["c3"]["capi_path"] = {k: r[k] for k in ("value", "unit", "ms_per_step", "e2e", "path", "gpu_launches")}
        except Exception as e:
            if rank == 0 and isinstance(extra.get("c3"), dict):
                extra["c3"]["capi_path"] = {"error": repr(e)[:300]}
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line["clocks"] = clocks
        line["numa"] = numa
        if extra:
            line["workloads"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline_for(w, 15.0)
            if cb:
                line["cpu_baseline"] = {"value": cb["value"], "unit": cb["unit"], "cores": cb["cores"], "kind": cb["kind"],
                                        "sample": cb["sample"], "extrapolated": cb.get("extrapolated")}
            if w["space"] == "coordinate" and args.workload == "c2":
                line["ref_cuda_baseline"] = ref_cuda_sample()
            if "c3" in extra and isinstance(extra["c3"], dict) and "error" not in extra["c3"]:
                extra["c3"]["ref_cuda_baseline"] = ref_cuda_sample(["--fft"])
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-worker":
        _worker_main(sys.argv[2:])
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--only", action="store_true", help="measure --workload only (no workloads.c3 / c4 side records)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--size", type=int, default=None, help="square frames of this edge instead of the workload's (sweeps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--fft-capi", action="store_true",
                    help="c3 through the reference-shaped C-ABI calls (autoenc_fft fft_l=1 + backprop_fft on real-space layers) "
                         "instead of the resident net's aefft_net_fft_step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        w = dict(WORKLOADS[args.workload])
        if args.size:
            w["Nx"] = w["Ny"] = args.size
        if args.batch is not None:
            w["batch"] = args.batch
        run_reference(args, w, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
