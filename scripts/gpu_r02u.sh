# round 2, twenty-first call (1 GPU): candidate pass over x-runs with the image-free loop; launch list; ncu of the filter and of the MODE 2 pair kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/u_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/u_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/u_ab.jsonl; }
run head
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zzz_fullsize.py -m gpu -q -p no:cacheprovider > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest_gpu.log
tail -3 gpurun_out/u_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/u_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nbr_filter -s 1 -c 1 -o gpurun_out/u_prof_k_nbr_filter python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/u_ncu_k_nbr_filter.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair2 -s 14 -c 1 -o gpurun_out/u_prof_k_pair2_mode2 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/u_ncu_k_pair2.log 2>&1
ls -la gpurun_out | tail -4
