# round 2, twenty-seventh call (1 GPU): candidate pass with one predicated chunk path; register cap of k_bonded again (the kernel changed twice since)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/aa_ab.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --kernels-only 2>gpurun_out/aa_$tag.err | grep '^{' | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/aa_ab.jsonl; }
run head
run bonded1 DDCB200_BONDED=1
run bonded8 DDCB200_BONDED=8
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_nbr -c 8 --csv --log-file gpurun_out/aa_nbr.csv python bench.py --steps 42 --warmup 3 --kernels-only > gpurun_out/aa_ncu.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "membership or cells or step0" > gpurun_out/aa_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/aa_pytest.log
tail -3 gpurun_out/aa_pytest.log
