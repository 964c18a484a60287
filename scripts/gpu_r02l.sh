# round 2, twelfth call (4 GPUs): halo overlapped (boundary rows on their own stream) vs inline (one launch), twice each
set -x
mkdir -p gpurun_out
rm -f gpurun_out/l_ab.jsonl
for rep in 1 2; do
for n in 4 2; do
  for h in overlap inline; do
    DDCB200_HALO=$h timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 20 --kernels-only 2>gpurun_out/l_${n}_$h.err | grep '^{' | sed "s/^{/{\"tag\": \"n${n}_${h}_$rep\", /" >> gpurun_out/l_ab.jsonl
  done
done
done
ls -la gpurun_out
