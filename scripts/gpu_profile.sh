# launch list + full ncu capture of one kernel (regex in $1, default k_pair); outputs under gpurun_out/
K=${1:-k_pair}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 22 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 2 -o gpurun_out/prof_$K -f python bench.py --steps 22 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_quick.json
