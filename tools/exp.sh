timeout -k 5 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "large or conv_fwd or concat" 2>&1 | tail -3
L="'k3 16 16 12 256 256' 'k3 32 16 12 256 256' 'k3 32 32 12 128 128' 'k3 64 32 12 128 128' 'k3 64 64 12 64 64' 'k3 16 32 12 256 256'"
for C in 2 3 4; do for E in 1 2; do echo "CTAS=$C EPI=$E"; eval CHAP_TC_CTAS=$C CHAP_TC_EPI=$E timeout -k 5 120 python tools/conv_bench.py $L; done; done
