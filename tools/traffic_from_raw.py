"""profiles/traffic.json from the raw metric tables of the `ncu --set full` captures (tools/profile_pass.sh):
per workload and per engine profile name the mean DRAM bytes (read + write) per launch, the DRAM and tensor-pipe utilisation.
bench.py fills roofline.traffic from this file.
usage: python tools/traffic_from_raw.py c2=gpurun_out/r3_c2_full_raw.csv c3=gpurun_out/r3_c3_full_raw.csv > profiles/traffic.json"""
import csv
import json
import sys

# kernel function name (substring) -> the name the engine's per-launch profile records use (ProfScope)
NAMES = [("wgrad_ts_kernel", "wgrad_ts"), ("conv_rs_kernel", "conv_rs"), ("gram_iter_bm_kernel", "spec_gram_iter"),
         ("gram_iter_ff_kernel", "spec_gram_iter"), ("gram_stats_bm_kernel", "spec_gram_stats"),
         ("gram_stats_ff_kernel", "spec_gram_stats"), ("spec_tc_kernel", "spec_contract_tc"), ("fft_rows_r2c", "fft_rows_r2c"),
         ("fft_rows_c2r", "fft_rows_c2r"), ("fft_cols", "fft_cols"), ("conv_reg_kernel", "spec_contract_reg"),
         ("kernel_spectrum_emb", "kernel_spectrum_emb"), ("binmajor_to_taps", "binmajor_to_taps"), ("gram_grad_kernel", "spec_gram_grad"),
         ("small_grad_kernel", "spec_small_grad"), ("small_mse_kernel", "spec_small_mse"), ("gradient_diff_tiled", "gradient_diff")]

out = {"_note": "mean over the captured launches of each kernel; DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum "
                "(ncu --set full --clock-control none, one step after warm-up; tools/profile_pass.sh, tools/traffic_from_raw.py)"}
for arg in sys.argv[1:]:
    wl, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[2:]:
        kn = r[ix["Kernel Name"]]
        name = next((p for sub, p in NAMES if sub in kn), None)
        if name is None:
            continue
        a = acc.setdefault(name, {"n": 0, "bytes": 0.0, "dram": 0.0, "tensor": 0.0, "ms": 0.0})
        unit_r, unit_w = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
        mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        a["bytes"] += float(r[ix["dram__bytes_read.sum"]]) * mul[unit_r] + float(r[ix["dram__bytes_write.sum"]]) * mul[unit_w]
        a["dram"] += float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])
        a["tensor"] += float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
        tu = rows[1][ix["gpu__time_duration.sum"]]
        a["ms"] += float(r[ix["gpu__time_duration.sum"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1e-6)
        a["n"] += 1
    out[wl] = {k: {"dram_bytes_per_launch": v["bytes"] / v["n"], "launches": v["n"], "dram_pct_of_peak_mean": v["dram"] / v["n"],
                   "tensor_pipe_pct_mean": v["tensor"] / v["n"], "ms_per_launch_under_ncu": v["ms"] / v["n"]} for k, v in acc.items()}
print(json.dumps(out, indent=1))
