# round 2, twenty-ninth call (1 GPU): final check of HEAD - the whole GPU suite, smoke, the driver's bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/ac_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ac_pytest_gpu.log
tail -4 gpurun_out/ac_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_smoke.log 2>&1; tail -1 gpurun_out/ac_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; tail -2 gpurun_out/ac_bench.err
