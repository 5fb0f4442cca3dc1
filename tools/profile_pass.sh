#!/bin/bash
# One profiling pass on a B200 (run under gpurun): per workload a plain run, the ncu launch list of the same command
# (gpu__time_duration only: cold-cache, serialised -> compare SHARES) and one `--set full` capture of the hot kernels.
# Outputs land in gpurun_out/<tag>_*; tools/ncu_summary.py turns the .ncu-rep files into the summaries under profiles/.
set -u
TAG=${1:-r2}
OUT=gpurun_out
for WL in c2 c3; do
  CMD="python bench.py --workload $WL --only --no-cpu-baseline --steps 2 --warmup 3"
  $CMD > $OUT/${TAG}_${WL}_plain.json 2> $OUT/${TAG}_${WL}_plain.err || { echo "plain run of $WL failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches_${WL}.csv $CMD \
      > $OUT/${TAG}_${WL}_ncu1.log 2>&1
  if [ $WL = c2 ]; then PAT='regex:wgrad_ts_kernel|conv_rs_kernel'; SKIP=36; CNT=12; else PAT='regex:spec_tc_kernel|gram_iter|gram_stats|fft_rows|fft_cols|conv_reg|kernel_spectrum_emb|binmajor_to_taps'; SKIP=99; CNT=33; fi
  ncu --set full --clock-control none -k "$PAT" -s $SKIP -c $CNT -o $OUT/${TAG}_${WL}_full $CMD \
      > $OUT/${TAG}_${WL}_ncu2.log 2>&1
  tail -2 $OUT/${TAG}_${WL}_ncu2.log
  # the reports are large (gpurun_out/ is capped at 64 MiB): keep the raw metric table and drop the report
  ncu -i $OUT/${TAG}_${WL}_full.ncu-rep --page raw --csv > $OUT/${TAG}_${WL}_full_raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/${TAG}_${WL}_full.ncu-rep "$WL: ncu --set full, one step after 3 warm-up steps ($CMD)" > $OUT/${TAG}_${WL}_ncu_summary.txt
  ls -la $OUT/${TAG}_${WL}_full.ncu-rep; rm -f $OUT/${TAG}_${WL}_full.ncu-rep
done
python tools/traffic_from_raw.py c2=$OUT/${TAG}_c2_full_raw.csv c3=$OUT/${TAG}_c3_full_raw.csv > $OUT/${TAG}_traffic.json
du -sh $OUT
