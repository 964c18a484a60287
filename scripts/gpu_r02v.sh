# round 2, twenty-second call (1 GPU): whole GPU suite at HEAD, kernels-only bench, the default bench line (e2e), launch list of the whole bench
set -x
mkdir -p gpurun_out
rm -f gpurun_out/v_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/v_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/v_ab.jsonl; }
run head
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest_gpu.log
tail -5 gpurun_out/v_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/v_launches_e2e.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-equilibration > gpurun_out/v_ncu_bench.log 2>&1
ls -la gpurun_out | tail -4
