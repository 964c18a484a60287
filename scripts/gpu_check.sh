set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 42 --warmup 3 --kernels-only > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair -s 5 -c 2 -o gpurun_out/prof_pair python bench.py --steps 22 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nbr_cell -c 1 -o gpurun_out/prof_nbr_cell env DDCB200_LISTBUILD=cell python bench.py --steps 3 --warmup 3 --kernels-only > gpurun_out/ncu_cell.log 2>&1
ls -la gpurun_out
