# round 2, eleventh call (4 GPUs): multi-rank parity (NGLF and NGLFCONSTRAINT) on the new slot order, scaling at 2 and 4 with the halo overlapped / inline
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_zzzzz_nglfc.py -m gpu -q -p no:cacheprovider -k "two_gpus or four_gpus or several_gpus" > gpurun_out/k_pytest_4gpu.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest_4gpu.log
for n in 2 4; do
  for h in overlap inline; do
    DDCB200_HALO=$h timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 20 --kernels-only 2>gpurun_out/k_k_${n}_$h.err | grep '^{' > gpurun_out/k_k_${n}_$h.json
  done
done
ls -la gpurun_out
