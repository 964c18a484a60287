# shortest possible evidence run: launch list, one full capture of k_pair, then a bench line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 42 --warmup 3 --kernels-only > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair -s 8 -c 1 -o gpurun_out/prof_k_pair -f python bench.py --steps 12 --warmup 3 --kernels-only > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat gpurun_out/bench_quick.json
