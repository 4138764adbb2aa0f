#!/bin/bash
# Round-2 ncu captures (one GPU): run each command without ncu first, then `--set full` on the conv kernels only (-c bounded).
set -u
mkdir -p gpurun_out
L2D="k3 128 128 12 32 32"; L3D="k3 64 64 2 28 28 20"; L0="k3 16 16 12 256 256"
python tools/conv_bench.py --reps 3 "$L0" "$L2D" "$L3D" "k3 128 128 2 14 14 10" "k3 256 256 12 16 16" > gpurun_out/r02_conv_bench.txt 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:"conv_tc_k|wgrad_tc_kernel" -c 3 -f -o gpurun_out/r02_conv_l0 python tools/conv_bench.py --reps 1 "$L0" > gpurun_out/ncu_l0.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"conv_tc_k|wgrad_tc_kernel" -c 3 -f -o gpurun_out/r02_conv_2d128 python tools/conv_bench.py --reps 1 "$L2D" > gpurun_out/ncu_2d128.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"conv_tc_k|wgrad_tc_kernel" -c 3 -f -o gpurun_out/r02_conv_3d64 python tools/conv_bench.py --reps 1 "$L3D" > gpurun_out/ncu_3d64.log 2>&1
CHAP_PRECISE_MAX_C=1048576 python tools/conv_bench.py --reps 3 "$L0" "$L2D" > gpurun_out/r02_conv_bench_precise.txt 2>&1
CHAP_PRECISE_MAX_C=1048576 ncu --set full --import-source on --clock-control none -k regex:"conv_tc_k" -c 2 -f -o gpurun_out/r02_conv_l0_precise python tools/conv_bench.py --reps 1 "$L0" > gpurun_out/ncu_l0p.log 2>&1
ls -la gpurun_out/*.ncu-rep
