# round 2, twenty-sixth call (1 GPU): candidate pass with bit masks; ncu of the pair launches over full and pruned rows and of the filter; launch list; GPU suite; the driver's bench line; smoke
set -x
mkdir -p gpurun_out
rm -f gpurun_out/z_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/z_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/z_ab.jsonl; }
run head
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest_gpu.log
tail -4 gpurun_out/z_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-equilibration > gpurun_out/z_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair2 -s 14 -c 2 -o gpurun_out/z_prof_k_pair2 python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/z_ncu_k_pair2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nbr_filter -s 1 -c 1 -o gpurun_out/z_prof_k_nbr_filter python bench.py --steps 22 --warmup 3 --kernels-only > gpurun_out/z_ncu_k_nbr_filter.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
ls -la gpurun_out | tail -3
