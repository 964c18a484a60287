# round 2, twenty-eighth call (8 GPUs): the driver's scaling line at 8 ranks at HEAD (strong scaling on 1M beads + the 10M-bead weak-scaling record)
set -x
mkdir -p gpurun_out
( time timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/ab_bench8.json 2> gpurun_out/ab_bench8.err ) 2>> gpurun_out/ab_bench8.err; echo "rc=$?" >> gpurun_out/ab_bench8.err
tail -3 gpurun_out/ab_bench8.err
