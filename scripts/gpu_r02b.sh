# round 2, second call (2 GPUs): the new decomposition on real NCCL + the shim failure under compute-sanitizer
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "two_gpus" > gpurun_out/b_pytest_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/b_bench2.json 2> gpurun_out/b_bench2.err; echo "rc=$?" >> gpurun_out/b_bench2.err
# the shim failure: ddcMD's NGLFCONSTRAINT (Langevin + constraints + barostat) on the host with the library as its potential
timeout 600 python - > gpurun_out/b_shim.log 2>&1 <<'PY'
import os, sys, subprocess, tempfile
sys.path.insert(0, "tests")
import test_zzzzzz_shim as ts
g = os.path.join(os.getcwd(), "tests", "golden")
tmp = tempfile.mkdtemp()
for exe, tag, args in ((ts.REF, "ref", []), (ts.SHIM, "gpu", ["1"]), (ts.SHIM_EMU, "emu", ["1"])):
    if not os.path.exists(exe):
        print("missing", exe); continue
    try:
        lines = ts.run_deck(g, "ras_small", tmp, exe, args, tag, "full")
        print("==", tag); print("\n".join(lines))
    except Exception as ex:
        print("==", tag, "FAILED", ex)
# again, under compute-sanitizer
d = os.path.join(tmp, "san"); os.makedirs(d)
import nglfc_decks, re
dk = nglfc_decks.make_variant(g, "ras_small", "full", d)
p = os.path.join(dk, "object.data"); s = open(p).read()
s = re.sub(r"deltaloop=\d+;", "deltaloop=6;", s); s = re.sub(r"printrate=\d+;", "printrate=1;", s); open(p, "w").write(s)
for tool in ("memcheck", "initcheck", "racecheck"):
    r = subprocess.run(["compute-sanitizer", "--tool", tool, "--print-limit", "20", ts.SHIM, "1"], cwd=dk, capture_output=True, text=True, timeout=500)
    print("== sanitizer", tool, "rc", r.returncode); print(r.stdout[-3000:]); print(r.stderr[-1500:])
PY
ls -la gpurun_out
