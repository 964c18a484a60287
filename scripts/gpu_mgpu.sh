# multi-GPU parity on N GPUs (default 2); output under gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
for d in popc_small ras_small waterbox; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_worker.py $d > gpurun_out/mgpu_$d.log 2>&1
  echo "$d rc=$?"; grep -E "step0 ok|MGPU_OK|Error|error|assert" gpurun_out/mgpu_$d.log | head -8
done
