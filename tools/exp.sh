timeout -k 5 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err || exit 1
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_r01_final.csv
