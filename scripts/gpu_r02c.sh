# round 2, third call (1 GPU): whole GPU suite again, k_pair2 variants A/B, list-build / bonded captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest_gpu.log
rm -f gpurun_out/c_pair_ab.jsonl
for v in old 1,1 1,12 2,1 2,8 2,10 3,8 4,1 4,8; do
  DDCB200_PAIR=$v timeout 200 python bench.py --steps 200 --warmup 20 --kernels-only 2>/dev/null | grep '^{' >> gpurun_out/c_pair_ab.jsonl
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 42 --warmup 3 --kernels-only > gpurun_out/c_ncu_bench.log 2>&1
for k in k_nbr_filter k_nbr_exact; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/c_prof_$k python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/c_ncu_$k.log 2>&1
done
for k in k_pair2 k_bonded; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/c_prof_$k python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/c_ncu_$k.log 2>&1
done
timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_bench.err
ls -la gpurun_out
