# round 2, eighth call (1 GPU): k_pair3 with prefetch, bonded register caps, row-ordering bin edges
set -x
mkdir -p gpurun_out
rm -f gpurun_out/h_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>/dev/null | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/h_ab.jsonl; }
run base       DDCB200_PAIR=2,8
run win        DDCB200_PAIR=win
run bonded1    DDCB200_BONDED=1
run bonded10   DDCB200_BONDED=10
run bonded12   DDCB200_BONDED=12
run edgesA     DDCB200_BIN_EDGES=0.0625,0.125,0.1875,0.25,0.375,0.5,0.6875
run edgesB     DDCB200_BIN_EDGES=0.05,0.1,0.15,0.2,0.275,0.375,0.55
run edgesC     DDCB200_BIN_EDGES=0.075,0.15,0.225,0.3,0.4,0.525,0.7
DDCB200_PAIR=win timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair3 -s 12 -c 1 -o gpurun_out/h_prof_k_pair3 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/h_ncu_k_pair3.log 2>&1
ls -la gpurun_out
