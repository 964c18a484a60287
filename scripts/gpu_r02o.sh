# round 2, fifteenth call (1 GPU): pruned rows A/B (every, margin), cache-policy hints of the pair walk, launch list, GPU suite with pruning on
set -x
mkdir -p gpurun_out
rm -f gpurun_out/o_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/o_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/o_ab.jsonl; }
run base DDCB200_PRUNE=0
run p4 DDCB200_PRUNE=4
run p5 DDCB200_PRUNE=5
run p4m12 DDCB200_PRUNE=4,1.2
run p5m15 DDCB200_PRUNE=5,1.5
run p10 DDCB200_PRUNE=10
run p2 DDCB200_PRUNE=2
run p4h1 DDCB200_PRUNE=4 DDCB200_PAIRHINT=1
run p4h2 DDCB200_PRUNE=4 DDCB200_PAIRHINT=2
run p4h4 DDCB200_PRUNE=4 DDCB200_PAIRHINT=4
run p4h5 DDCB200_PRUNE=4 DDCB200_PAIRHINT=5
run h1 DDCB200_PRUNE=0 DDCB200_PAIRHINT=1
DDCB200_PRUNE=4 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/o_pytest_gpu_prune4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest_gpu_prune4.log
tail -5 gpurun_out/o_pytest_gpu_prune4.log
DDCB200_PRUNE=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/o_launches.csv python bench.py --steps 42 --warmup 3 --kernels-only --no-equilibration > gpurun_out/o_ncu_bench.log 2>&1
DDCB200_PRUNE=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair2 -s 13 -c 1 -o gpurun_out/o_prof_k_pair2_mode2 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/o_ncu_k_pair2.log 2>&1
ls -la gpurun_out | tail -8
