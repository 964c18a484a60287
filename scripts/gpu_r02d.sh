# round 2, fourth call (1 GPU): k_pair2 variants A/B, captures of k_pair2 / k_bonded, parity subset on the default variant
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzz_fullsize.py -m gpu -q -p no:cacheprovider > gpurun_out/d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest_gpu.log
rm -f gpurun_out/d_pair_ab.jsonl
for v in old 1,1 1,12 2,1 2,8 2,10 3,8 4,1 4,8; do
  DDCB200_PAIR=$v timeout 200 python bench.py --steps 200 --warmup 20 --kernels-only 2>/dev/null | grep '^{' >> gpurun_out/d_pair_ab.jsonl
done
for k in k_pair2 k_bonded; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/d_prof_$k python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/d_ncu_$k.log 2>&1
done
DDCB200_PAIR=1,1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair2 -s 12 -c 1 -o gpurun_out/d_prof_k_pair2_11 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/d_ncu_k_pair2_11.log 2>&1
ls -la gpurun_out
