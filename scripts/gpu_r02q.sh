# round 2, seventeenth call (1 GPU): one-pass list build (k_nbr_build) against the two passes, with and without pruned rows; GPU suite
set -x
mkdir -p gpurun_out
rm -f gpurun_out/q_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/q_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/q_ab.jsonl; }
run twopass_p0 DDCB200_LISTBUILD=twopass DDCB200_PRUNE=0
run fused_p0 DDCB200_LISTBUILD=fused DDCB200_PRUNE=0
run twopass_p4 DDCB200_LISTBUILD=twopass DDCB200_PRUNE=4
run fused_p4 DDCB200_LISTBUILD=fused DDCB200_PRUNE=4
run fused_p5 DDCB200_LISTBUILD=fused DDCB200_PRUNE=5
DDCB200_PRUNE=4 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest_gpu.log
tail -5 gpurun_out/q_pytest_gpu.log
DDCB200_PRUNE=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/q_ncu_bench.log 2>&1
DDCB200_PRUNE=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nbr_build -s 1 -c 1 -o gpurun_out/q_prof_k_nbr_build python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/q_ncu_k_nbr_build.log 2>&1
ls -la gpurun_out | tail -4
