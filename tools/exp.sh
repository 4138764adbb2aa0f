L='k3 16 16 12 256 256'
for D in 0 1 2 3; do echo "WG_DEBUG=$D"; CHAP_WG_DEBUG=$D timeout -k 5 120 python tools/conv_bench.py "$L" 'k3 32 32 12 128 128'; done
