#!/bin/bash
# ncu evidence for the round: full capture of the dataflow frame kernel (second launch: warm) and the launch list of the
# bench command.  Only the small CSV exports are kept (gpurun_out/ is limited to 64 MiB).
set -u
mkdir -p gpurun_out
O=gpurun_out
R=/tmp/r1_mega2_full
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_frames_mega2 -s 1 -c 1 -f -o $R \
  python tools/profile_frame.py --frames 48 > $O/r1_mega2_full.log 2>&1
ncu -i $R.ncu-rep --page raw --csv > $O/r1_mega2_full.raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r1_launches_bench_final.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r1_launches_bench_final.log 2>&1
ls -la $O | tail -6
