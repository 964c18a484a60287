# round 2, thirtieth call (2 GPUs): the driver's scaling line at 2 ranks with the clock sampler started before the warm-up
set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/ad_bench2.json 2> gpurun_out/ad_bench2.err; echo "rc=$?" >> gpurun_out/ad_bench2.err
tail -3 gpurun_out/ad_bench2.err
