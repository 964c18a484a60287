# round 2, thirty-first call (1 GPU): last sanity check of HEAD - smoke, parity and variant tests
set -x
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ae_smoke.log 2>&1; tail -1 gpurun_out/ae_smoke.log
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_zzzzzzz_variants.py -m gpu -q -p no:cacheprovider > gpurun_out/ae_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ae_pytest.log
tail -3 gpurun_out/ae_pytest.log
