export CHAP_B200_LIB=$PWD/chap_b200/lib/libchap_b200_dbg.so
for D in 64 128; do echo "DEBUG=$D"; CHAP_TC_DEBUG=$D timeout -k 5 120 python tools/tc_trace.py 'k3 128 128 12 32 32' 'k1 256 128 12 16 16'; done
