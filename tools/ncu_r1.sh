#!/bin/bash
# Round-1 ncu evidence.  Run on the GPU box:  bash tools/ncu_r1.sh   (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. launch list of the bench command (serialised, cold-cache per-launch times: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r1_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r1_launches_bench.log 2>&1
# 2. full captures: the persistent decode kernel (second launch: warm), one big vocoder conv, the tcgen05 GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_frames_mega -s 1 -c 1 -f -o $O/r1_mega_full \
  python tools/profile_frame.py --frames 48 > $O/r1_mega_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 2 -f -o $O/r1_gemm_tc_full \
  python tools/profile_frame.py --frames 2 > $O/r1_gemm_tc_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:voc_conv_mma_kernel -s 30 -c 8 -f -o $O/r1_voc_mma_full \
  python tools/profile_frame.py --frames 64 --vocoder > $O/r1_voc_mma_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"fused_residual_rmsnorm_large|sample_kernel" -c 9 -f -o $O/r1_ops_full \
  python tools/profile_ops.py > $O/r1_ops_full.log 2>&1
for n in r1_mega_full r1_gemm_tc_full r1_voc_mma_full r1_ops_full; do
  ncu -i $O/$n.ncu-rep --page raw --csv > $O/$n.raw.csv 2>/dev/null
done
ls -la $O
