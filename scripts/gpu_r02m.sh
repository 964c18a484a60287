# round 2, thirteenth call (8 GPUs): 8-rank parity, the driver's scaling line at 8 (with the 10M weak-scaling record), per-kernel times
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt; nproc >> gpurun_out/m_smi.txt; free -g >> gpurun_out/m_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "eight_gpus" > gpurun_out/m_pytest_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/m_pytest_8gpu.log
for h in overlap inline; do
  DDCB200_HALO=$h timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 200 --warmup 20 --kernels-only 2>gpurun_out/m_k_8_$h.err | grep '^{' > gpurun_out/m_k_8_$h.json
done
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/m_bench8.json 2> gpurun_out/m_bench8.err ) 2>> gpurun_out/m_bench8.err; echo "rc=$?" >> gpurun_out/m_bench8.err
ls -la gpurun_out
