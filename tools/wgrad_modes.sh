L="k3 32 16 12 256 256|k3 32 32 12 128 128|k3 64 32 12 128 128|k3 64 64 12 64 64|k3 128 64 12 64 64|k3 32 32 2 56 56 40|k3 64 64 2 28 28 20"
IFS='|' read -ra LL <<< "$L"
for mode in "X=1" "CHAP_WG_NO_V4=1"; do
echo "== $mode"
env $mode python tools/conv_bench.py --reps 5 "${LL[@]}" 2>&1 | grep "^k3" | sed 's/in .*conv_tc_fwd/ fwd/' 
done
