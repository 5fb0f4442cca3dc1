#!/bin/bash
# Profiling pass for the workloads outside the default line's ncu evidence (run under gpurun, one GPU): c4 (5 pairs, FFT
# space, multiobjective term; one-GPU form of the bin-sharded step) and c1 (CPU `backprop` semantics on the GPU): a plain
# run, the ncu launch list of the same command, and one `--set full` capture of the tensor-core multiobjective kernel.
set -u
TAG=${1:-r2}
OUT=gpurun_out
for WL in c4 c1; do
  CMD="python bench.py --workload $WL --only --no-cpu-baseline --steps 2 --warmup 3"
  timeout 300 $CMD > $OUT/${TAG}_${WL}_plain.json 2> $OUT/${TAG}_${WL}_plain.err || { echo "plain run of $WL failed"; continue; }
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches_${WL}.csv $CMD \
      > $OUT/${TAG}_${WL}_ncu1.log 2>&1
done
CMD="python tools/gdiff_probe.py"
timeout 120 $CMD > $OUT/${TAG}_gdiff_plain.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gdiff_tc_kernel -s 2 -c 2 -o $OUT/${TAG}_gdiff_full $CMD > $OUT/${TAG}_gdiff_ncu2.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_gdiff_full.ncu-rep "gdiff_tc_kernel at 128 -> 256 channels (32 768 kernels per tensor): ncu --set full ($CMD)" > $OUT/${TAG}_gdiff_ncu_summary.txt
ncu -i $OUT/${TAG}_gdiff_full.ncu-rep --page raw --csv > $OUT/${TAG}_gdiff_full_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_gdiff_full.ncu-rep
du -sh $OUT
