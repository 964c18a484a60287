# round 2, sixteenth call (1 GPU): pruned rows, <every> x <margin as a fraction of deltaR>; rows loaded with L1::no_allocate
set -x
mkdir -p gpurun_out
rm -f gpurun_out/p_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/p_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/p_ab.jsonl; }
run base DDCB200_PRUNE=0
run p4f25 DDCB200_PRUNE=4,0.25
run p4f30 DDCB200_PRUNE=4,0.30
run p4f35 DDCB200_PRUNE=4,0.35
run p5f30 DDCB200_PRUNE=5,0.30
run p5f35 DDCB200_PRUNE=5,0.35
run p5f40 DDCB200_PRUNE=5,0.40
run p3f20 DDCB200_PRUNE=3,0.20
run p3f25 DDCB200_PRUNE=3,0.25
run p7f50 DDCB200_PRUNE=7,0.50
run p10f65 DDCB200_PRUNE=10,0.65
DDCB200_PRUNE=4,0.3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 22 --warmup 3 --kernels-only --no-equilibration > gpurun_out/p_ncu_bench.log 2>&1
ls -la gpurun_out | tail -4
