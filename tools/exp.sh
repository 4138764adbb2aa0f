timeout -k 5 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -3
timeout -k 5 120 python tools/conv_bench.py 'k3 16 4 12 256 256'
CHAP_NO_HEAD_TC=1 timeout -k 5 120 python tools/conv_bench.py 'k3 16 4 12 256 256'
