#!/bin/bash
# Final evidence run of round 2 (one B200): bench lines, launch list of the bench command, sanitizers, all configurations.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_bench_final.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-by-batch > $O/r2_launches_bench_final.log 2>&1
rm -f $O/r2_sanitizer_final_summary.txt
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_tiny.py tiny > $O/r2_sanitizer_final_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" >> $O/r2_sanitizer_final_summary.txt
  tail -4 $O/r2_sanitizer_final_$tool.log >> $O/r2_sanitizer_final_summary.txt
done
timeout 900 python tools/bench_configs.py > $O/r2_baseline_configs_final.txt 2>&1
tail -c 600 $O/r2_bench_final.json; echo; tail -c 400 $O/r2_bench_reference_arm.json; echo
cat $O/r2_sanitizer_final_summary.txt; tail -12 $O/r2_baseline_configs_final.txt
