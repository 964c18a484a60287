# round 2, fifth call (1 GPU): whole GPU suite (with the 1M live-reference test), the new bench line, the reference arm on the real deck
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err ) 2>> gpurun_out/e_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bonded -s 12 -c 1 -o gpurun_out/e_prof_k_bonded python bench.py --steps 22 --warmup 3 --kernels-only --no-equilibration > gpurun_out/e_ncu_k_bonded.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only > gpurun_out/e_kernels.json 2>/dev/null
ls -la gpurun_out
