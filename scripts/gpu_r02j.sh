# round 2, tenth call (1 GPU): neighbourhood walk bound A/B, whole GPU suite
set -x
mkdir -p gpurun_out
rm -f gpurun_out/j_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>/dev/null | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/j_ab.jsonl; }
run walk_cell DDCB200_WALK=cell
run walk_bead DDCB200_WALK=bead
run walk_global DDCB200_WALK=global
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair2 -s 12 -c 1 -o gpurun_out/j_prof_k_pair2 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/j_ncu_k_pair2.log 2>&1
ls -la gpurun_out
