# round 2, twenty-fourth call (1 GPU): far-segment entries parked and worked off together (this build) against the plain walk (libddcmd_b200_nodefer.so),
# CTA-to-tile order for L1 sharing (DDCB200_TILEORDER=zy); default bench line with the capped energy kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/x_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/x_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/x_ab.jsonl; }
run nodefer DDCB200_LIBFILE=libddcmd_b200_nodefer.so
run defer
run nodefer_zy DDCB200_LIBFILE=libddcmd_b200_nodefer.so DDCB200_TILEORDER=zy
run defer_zy DDCB200_TILEORDER=zy
DDCB200_LIBFILE=libddcmd_b200_nodefer.so DDCB200_TILEORDER=zy timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair2 -s 14 -c 1 -o gpurun_out/x_prof_k_pair2_mode2_zy python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/x_ncu_k_pair2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zzzzzzz_variants.py -m gpu -q -p no:cacheprovider > gpurun_out/x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest_gpu.log
tail -3 gpurun_out/x_pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
ls -la gpurun_out | tail -3
