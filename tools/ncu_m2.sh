#!/bin/bash
# ncu full capture of the dataflow frame kernel (second launch: warm), with source-level sampling
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_frames_mega2 -s 1 -c 1 -f -o $O/r1_mega2_full \
  python tools/profile_frame.py --frames 48 > $O/r1_mega2_full.log 2>&1
ncu -i $O/r1_mega2_full.ncu-rep --page raw --csv > $O/r1_mega2_full.raw.csv 2>/dev/null
ncu -i $O/r1_mega2_full.ncu-rep --page source --csv > $O/r1_mega2_full.source.csv 2>/dev/null
ls -la $O
