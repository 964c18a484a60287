# round 2, twentieth call (1 GPU): one-sweep exact pass with two-segment rows (k_nbr_exact2) against the bin-sorting one; bonded sum inside the pair kernel; GPU suite
set -x
mkdir -p gpurun_out
rm -f gpurun_out/t_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/t_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/t_ab.jsonl; }
run bins DDCB200_LISTBUILD=bins
run twoseg DDCB200_LISTBUILD=twoseg
run twoseg_n40 DDCB200_LISTBUILD=twoseg DDCB200_NEAR=0.4
run twoseg_p0 DDCB200_LISTBUILD=twoseg DDCB200_PRUNE=0
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest_gpu.log
tail -5 gpurun_out/t_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/t_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nbr_exact2 -s 1 -c 1 -o gpurun_out/t_prof_k_nbr_exact2 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/t_ncu_k_nbr_exact2.log 2>&1
ls -la gpurun_out | tail -4
