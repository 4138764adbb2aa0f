# A/B of library builds on the same GPU: bash tools/ab_libs.sh lib1.so lib2.so ...   (env NBUF passes CHAP_TC_NBUF)
for i in 1 2; do
for lib in "$@"; do
CHAP_B200_LIB=$PWD/$lib python bench.py --steps 40 --warmup 3 --no-extras --no-cpu-baseline --no-kernel-timing 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['ms_per_step'])"
done; done
