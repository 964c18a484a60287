# round 2, eighteenth call (1 GPU): one-pass tile build (k_nbr_tile, rows in two segments) against the two passes; GPU suite; ncu
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/r_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/r_ab.jsonl; }
run twopass_p4 DDCB200_LISTBUILD=twopass DDCB200_PRUNE=4
run tile_p4 DDCB200_LISTBUILD=fused DDCB200_PRUNE=4
run tile_p0 DDCB200_LISTBUILD=fused DDCB200_PRUNE=0
run tile_p5 DDCB200_LISTBUILD=fused DDCB200_PRUNE=5,0.3
run tile_p4n40 DDCB200_LISTBUILD=fused DDCB200_PRUNE=4 DDCB200_NEAR=0.4
run tile_p4n28 DDCB200_LISTBUILD=fused DDCB200_PRUNE=4 DDCB200_NEAR=0.28
DDCB200_PRUNE=4 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest_gpu.log
tail -5 gpurun_out/r_pytest_gpu.log
DDCB200_PRUNE=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/r_ncu_bench.log 2>&1
DDCB200_PRUNE=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nbr_tile -s 1 -c 1 -o gpurun_out/r_prof_k_nbr_tile python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/r_ncu_k_nbr_tile.log 2>&1
ls -la gpurun_out | tail -4
