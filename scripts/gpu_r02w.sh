# round 2, twenty-third call (1 GPU): the end-to-end figure and its parts, register caps of the energy instantiation of the pair kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/w_e2e.jsonl
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/w_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/w_e2e.jsonl; }
run e1
run e6 DDCB200_PAIR=2,8,6
run e5 DDCB200_PAIR=2,8,5
run pf1 DDCB200_PAIR=1,1
run e1_p0 DDCB200_PRUNE=0
ls -la gpurun_out | tail -3
