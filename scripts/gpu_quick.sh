# quick GPU check: parity tests + a short bench (no CPU baseline); output under gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?" >> gpurun_out/bench_quick.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_quick.json
