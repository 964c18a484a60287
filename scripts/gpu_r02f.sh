# round 2, sixth call (1 GPU): windowed pair kernel vs k_pair2, bonded register cap, the fixed scan; parity on the windowed path
set -x
mkdir -p gpurun_out
DDCB200_PAIR=win timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzz_fullsize.py -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest_win.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest_win.log
rm -f gpurun_out/f_ab.jsonl
for v in 2,8 win; do
  for b in plain capped; do
    DDCB200_PAIR=$v DDCB200_BONDED=$b timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>/dev/null | grep '^{' | sed "s/^{/{\"bonded\": \"$b\", /" >> gpurun_out/f_ab.jsonl
  done
done
DDCB200_PAIR=win timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair3 -s 12 -c 1 -o gpurun_out/f_prof_k_pair3 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/f_ncu_k_pair3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 42 --warmup 3 --kernels-only --no-equilibration > gpurun_out/f_ncu_bench.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
ls -la gpurun_out
