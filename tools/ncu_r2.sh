#!/bin/bash
# ncu / sanitizer evidence of round 2 (one B200).  Only small exports are kept (gpurun_out/ is limited to 64 MiB).
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. full capture of the dataflow frame kernel of the PRODUCT build (second launch: warm), one launch = 16 frames
R=/tmp/r2_mega2_full
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_frames_mega2 -s 1 -c 1 -f -o $R \
  python tools/profile_frame.py --frames 48 > $O/r2_mega2_full.log 2>&1
ncu -i $R.ncu-rep --page raw --csv > $O/r2_mega2_full.raw.csv 2>/dev/null
# 2. the same for the TMA-ring generation (opt-in)
R4=/tmp/r2_mega4_full
Q3_MEGA=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_frames_mega4 -s 1 -c 1 -f -o $R4 \
  python tools/profile_frame.py --frames 48 > $O/r2_mega4_full.log 2>&1
ncu -i $R4.ncu-rep --page raw --csv > $O/r2_mega4_full.raw.csv 2>/dev/null
# 3. launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-by-batch > $O/r2_launches_bench.log 2>&1
# 4. compute-sanitizer on the tiny model (persistent kernel + sampler + vocoder)
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_tiny.py tiny > $O/r2_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" >> $O/r2_sanitizer_summary.txt
  tail -4 $O/r2_sanitizer_$tool.log >> $O/r2_sanitizer_summary.txt
done
ls -la $O | tail -12
cat $O/r2_sanitizer_summary.txt
