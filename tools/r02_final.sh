#!/bin/bash
# Round-2 final evidence on ONE GPU: the default bench line, the reference arm, the ncu launch list of one replayed iteration and
# `--set full` captures of the three reworked kernels.  Every ncu command runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -5 gpurun_out/r02_bench_1gpu.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
B="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-kernel-timing"
$B > gpurun_out/r02_launch_run.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4500 -c 2600 --csv --log-file gpurun_out/r02_launches_2d.csv $B > gpurun_out/ncu_launch.log 2>&1
L0="k3 16 16 12 256 256"; L1="k3 32 32 12 128 128"; L3="k3 16 16 2 112 112 80"
python tools/conv_bench.py --reps 3 "$L0" "$L1" "$L3" "k3 32 16 2 112 112 80" "k3 64 32 12 128 128" > gpurun_out/r02_conv_bench_final.txt 2>&1 || exit 1
for tag in l0 l1 l3; do
  case $tag in l0) L="$L0";; l1) L="$L1";; l3) L="$L3";; esac
  ncu --set full --import-source on --clock-control none -k regex:"conv_tc_k|wgrad_tc_kernel" -c 3 -f -o gpurun_out/r02f_conv_$tag python tools/conv_bench.py --reps 1 "$L" > gpurun_out/ncu_f_$tag.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
