# round 2, first call: the whole GPU suite (no -x: every failure is wanted), bench, launch list, full captures of what ships
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt; free -g >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 42 --warmup 3 --kernels-only > gpurun_out/ncu_bench.log 2>&1
for k in k_pair k_nbr_filter k_nbr_exact k_bonded k_integrate; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
