"""BASELINE configs[4]: batch / resolution sweep, coordinate vs FFT space, on the GPUs of this box.  Runs bench.py once per
point (resident frames; --no-cpu-baseline) and prints a markdown table (kept under profiles/).
usage: python tools/sweep.py [--gpus N] > profiles/r2_sweep.md"""
import argparse, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
points = [("c2", 256, 64), ("c2", 512, 64), ("c2", 1024, 16), ("c2", 2048, 4), ("c2", 4096, 1),
          ("c3", 256, 128), ("c3", 512, 128), ("c3", 1024, 64), ("c3", 2048, 16), ("c3", 4096, 4)]
if a.quick:
    points = [p for p in points if p[1] <= 512]
print("| workload | frame | frames per GPU | GPUs | frames/s (resident) | ms/step | frames/s (end to end, fp32 frames) | top kernel (share) |")
print("|---|---|---|---|---|---|---|---|")
for wl, size, batch in points:
    cmd = [sys.executable]
    if a.gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
                "--master-port", "29533"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--workload", wl, "--size", str(size), "--batch", str(batch),
            "--steps", "3", "--warmup", "3", "--no-cpu-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = next((l for l in reversed(r.stdout.strip().splitlines()) if l.startswith("{")), None)
    if not line:
        print(f"| {wl} | {size}x{size} | {batch} | {a.gpus} | failed: {r.stderr.strip()[-120:]} | | | |")
        continue
    d = json.loads(line)
    rf = d.get("roofline") or {}
    print(f"| {wl} | {size}x{size} | {batch} | {a.gpus} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['e2e']['value']:.0f} | "
          f"{rf.get('kernel')} ({100 * (rf.get('share_of_step') or 0):.0f}%) |", flush=True)
