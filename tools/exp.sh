L="'k3 128 128 12 32 32' 'k3 64 64 12 64 64' 'k3 256 128 12 32 32' 'k3 256 256 12 16 16' 'k3 64 32 12 128 128' 'k3 128 64 12 64 64'"
for B in 1000000 2000000 4000000; do echo "BUDGET=$B"; eval CHAP_WG_BUDGET=$B timeout -k 5 120 python tools/conv_bench.py $L; done
