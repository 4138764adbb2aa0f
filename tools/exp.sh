timeout -k 5 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "large or conv_fwd" 2>&1 | tail -5
timeout -k 5 120 python tools/conv_bench.py 'up2 32 16 12 128 128' 'up2 64 32 12 64 64' 'up2 128 64 12 32 32' 'up2 256 128 12 16 16' 'k1 32 16 12 128 128' 'k1 256 128 12 16 16'
