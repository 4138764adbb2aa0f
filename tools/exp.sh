L="'k3 16 16 12 256 256' 'k3 16 32 12 256 256' 'k3 32 16 12 256 256' 'k3 16 4 12 256 256' 'k3 128 128 12 32 32'"
eval timeout -k 5 120 python tools/conv_bench.py $L
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -k 5 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h10.json 2> gpurun_out/bench_h10.err; tail -2 gpurun_out/bench_h10.err
