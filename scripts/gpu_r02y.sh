# round 2, twenty-fifth call (4 GPUs): 2- and 4-rank parity tests (NGLF and NGLFCONSTRAINT) at HEAD, strong-scaling lines at 4 and 2 ranks
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/y_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzzzz_nglfc.py -m gpu -q -p no:cacheprovider -k "two_gpus or four_gpus or several_gpus" > gpurun_out/y_pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_mgpu.log
tail -4 gpurun_out/y_pytest_mgpu.log
rm -f gpurun_out/y_ab.jsonl
for n in 4 2; do
  for h in overlap inline; do
    DDCB200_HALO=$h timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 200 --warmup 20 --kernels-only 2>gpurun_out/y_${n}_$h.err | grep '^{' | sed "s/^{/{\"tag\": \"n${n}_${h}\", /" >> gpurun_out/y_ab.jsonl
  done
done
ls -la gpurun_out | tail -3
