# round 2, seventh call (4 GPUs): 4-rank parity on real NCCL, strong scaling at 4 with the halo inline / overlapped
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g_smi.txt; nvidia-smi topo -m >> gpurun_out/g_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "two_gpus or four_gpus" > gpurun_out/g_pytest_4gpu.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest_4gpu.log
for n in 1 2 4; do
  for h in overlap inline; do
    if [ $n = 1 ] && [ $h = inline ]; then continue; fi
    if [ $n = 1 ]; then
      DDCB200_HALO=$h timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 --kernels-only 2>gpurun_out/g_k_${n}_$h.err | grep '^{' > gpurun_out/g_k_${n}_$h.json
    else
      DDCB200_HALO=$h timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 200 --warmup 20 --kernels-only 2>gpurun_out/g_k_${n}_$h.err | grep '^{' > gpurun_out/g_k_${n}_$h.json
    fi
  done
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/g_bench4.json 2> gpurun_out/g_bench4.err; echo "rc=$?" >> gpurun_out/g_bench4.err
ls -la gpurun_out
