"""Per-kernel roofline table from bench.py JSON lines (their `kernels` lists carry the ALGORITHMIC TFLOP/s and GB/s of every
kernel, measured with CUDA events inside the timed region): fraction of the measured HBM peak for every kernel, and of the
sustained dense-bf16 rate for the tensor-core kernels (x3 for the BF16X3 / 3xTF32 passes = share of the tensor pipe's time).
usage: python tools/roofline_table.py profiles/r2_bench_n1.json [profiles/r2_bench_c4_n1.json ...] > profiles/r2_roofline_by_kernel.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM, TF = pk["hbm_gbs"], pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
TENSOR = ("_rs", "_ts", "_tc", "gradient_diff")


def table(name, rec):
    print(f"\n## {name}: {rec['ms_per_step']:.3f} ms per step, {rec['value']:.0f} {rec['unit']} ({rec['n_gpus']} GPU)\n")
    print("| kernel | launches / step | ms / step | share | GB/s (algorithmic) | of HBM peak | TFLOP/s (algorithmic) | of sustained bf16 (x passes) |")
    print("|---|---|---|---|---|---|---|---|")
    tot = sum(k["ms_per_step"] for k in rec["kernels"])
    for k in rec["kernels"]:
        tc = k["name"].endswith(TENSOR[:3]) or k["name"] == TENSOR[3]
        tfrac = f"{k['tflops'] / TF:.3f} ({3 * k['tflops'] / TF:.2f})" if tc and k["tflops"] > 0 else "—"
        print(f"| `{k['name']}` | {k['launches_per_step']:.0f} | {k['ms_per_step']:.3f} | {100 * k['ms_per_step'] / rec['ms_per_step_instrumented'] if 'ms_per_step_instrumented' in rec else 100 * k['ms_per_step'] / tot:.1f} % | "
              f"{k['gbs']:.0f} | {k['gbs'] / HBM:.2f} | {k['tflops']:.1f} | {tfrac} |")


print("# Per-kernel roofline fractions (round 2, final code)\n")
print(f"Peaks: `MEASURED_PEAKS.json` — HBM {HBM:.0f} GB/s, dense bf16 sustained {TF:.0f} TFLOP/s.  Kernel times are CUDA-event times inside "
      "the timed region of `bench.py` (the instrumented pass); GB/s and TFLOP/s are the algorithmic bytes / FLOP of SURVEY 8(d) over those "
      "times.  A tensor-core kernel computing fp32-grade results runs three MMA passes (BF16X3 / 3xTF32): the bracketed figure is its share "
      "of the tensor pipe's time.  Produced by `tools/roofline_table.py`.")
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    table(d["config"]["workload"].split(":")[0], d)
    for nm, w in (d.get("workloads") or {}).items():
        if isinstance(w, dict) and "kernels" in w:
            table(nm, w)
