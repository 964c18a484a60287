# round 2, fourteenth call (1 GPU): term-once k_bonded (bench + whole GPU suite), gather microbenchmark
set -x
mkdir -p gpurun_out
rm -f gpurun_out/n_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/n_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/n_ab.jsonl; }
run base DDCB200_BONDED=12
run bonded8 DDCB200_BONDED=8
run bonded1 DDCB200_BONDED=1
timeout 120 scripts/microbench/gather > gpurun_out/n_gather.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest_gpu.log
tail -5 gpurun_out/n_pytest_gpu.log
cat gpurun_out/n_gather.txt
