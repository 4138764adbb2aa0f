timeout -k 5 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "large or conv_fwd or concat" 2>&1 | tail -2
CHAP_B200_LIB=$PWD/chap_b200/lib/libchap_b200_dbg.so timeout -k 5 120 python tools/tc_trace.py 'k3 128 128 12 32 32' 'k1 256 128 12 16 16' 'k3 16 16 12 256 256'
L="'k3 16 16 12 256 256' 'k3 32 16 12 256 256' 'k3 64 32 12 128 128' 'k3 128 128 12 32 32' 'k3 256 128 12 32 32'"
eval timeout -k 5 120 python tools/conv_bench.py $L
