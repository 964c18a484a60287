# round 2, nineteenth call (1 GPU): k_nbr_tile with the image-free scan loop and stencil culling against the two passes; GPU suite; ncu
set -x
mkdir -p gpurun_out
rm -f gpurun_out/s_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/s_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/s_ab.jsonl; }
run twopass DDCB200_LISTBUILD=twopass
run tile DDCB200_LISTBUILD=fused
run tile_n40 DDCB200_LISTBUILD=fused DDCB200_NEAR=0.4
run tile_n50 DDCB200_LISTBUILD=fused DDCB200_NEAR=0.5
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest_gpu.log
tail -5 gpurun_out/s_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s_launches.csv python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/s_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nbr_tile -s 1 -c 1 -o gpurun_out/s_prof_k_nbr_tile python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/s_ncu_k_nbr_tile.log 2>&1
ls -la gpurun_out | tail -4
