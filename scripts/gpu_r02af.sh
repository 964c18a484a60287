# round 2, thirty-second call (1 GPU): margin of the pruned rows once more, with the two-segment rows (near edge moved along)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/af_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 100 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/af_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/af_ab.jsonl; }
run p4m28
run p4m33 DDCB200_PRUNE=4,0.33 DDCB200_NEAR=0.35
run p5m38 DDCB200_PRUNE=5,0.38 DDCB200_NEAR=0.40
run p3m24 DDCB200_PRUNE=3,0.24
