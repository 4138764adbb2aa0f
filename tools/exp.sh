L='k3 16 16 12 256 256'
timeout -k 5 120 python tools/conv_bench.py --reps 1 "$L" || exit 1
timeout -k 5 300 ncu --set full --import-source on --clock-control none -k regex:"conv_tc_kernel|wgrad_tc_kernel" -c 3 -f -o gpurun_out/r01_conv_final python tools/conv_bench.py --reps 1 "$L" > gpurun_out/ncu_convfinal.log 2>&1
tail -2 gpurun_out/ncu_convfinal.log
