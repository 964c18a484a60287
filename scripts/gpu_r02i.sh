# round 2, ninth call (1 GPU): per-cell candidate pass vs per-bead, the whole GPU suite on the new slot order
set -x
mkdir -p gpurun_out
rm -f gpurun_out/i_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>/dev/null | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/i_ab.jsonl; }
run cell DDCB200_FILTER=cell
run bead DDCB200_FILTER=bead
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 42 --warmup 3 --kernels-only --no-equilibration > gpurun_out/i_ncu_bench.log 2>&1
for k in k_nbr_filter_cell k_nbr_exact; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/i_prof_$k python bench.py --steps 22 --warmup 3 --kernels-only --no-equilibration > gpurun_out/i_ncu_$k.log 2>&1
done
ls -la gpurun_out
