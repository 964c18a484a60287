#!/bin/bash
# A/B runs for the knobs that were added without a GPU at hand (round 1, sessions 8+): one gpurun call, JSON lines under
# gpurun_out/.  Usage:  gpurun --timeout 1500 -- 'bash scripts/ab_round2.sh'
#   list build:  two-pass vs one-pass cell build (bit-identical rows; "auto" times both and keeps the faster)
#   row order:   DDCB200_BIN_EDGES - 8 bins (default) vs 4 / 2 / 1 bins: fewer bins = rows closer to slot order
#                (fewer distinct 128-byte lines per gather step, scripts/pair_locality_stats.py) but a coarser trimmed walk
set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab_round2.jsonl
: > "$OUT"
run() {  # name, env...
    local name="$1"; shift
    echo "== $name" >&2
    env "$@" python bench.py --steps 120 --warmup 20 --no-cpu-baseline 2> "gpurun_out/ab_$name.log" | sed "s/^{/{\"variant\": \"$name\", /" >> "$OUT"
}
run auto            DDCB200_LISTBUILD=auto
run twopass         DDCB200_LISTBUILD=twopass
run cell            DDCB200_LISTBUILD=cell
run cell_walkglobal DDCB200_LISTBUILD=cell DDCB200_WALK=global
run cell_bins4      DDCB200_LISTBUILD=cell DDCB200_BIN_EDGES=-0.25,-0.25,0.0,0.0,0.25,0.25,0.625
run cell_bins2      DDCB200_LISTBUILD=cell DDCB200_BIN_EDGES=-0.25,-0.25,-0.25,0.25,0.25,0.25,0.25
run cell_bins1      DDCB200_LISTBUILD=cell DDCB200_BIN_EDGES=1.0,1.0,1.0,1.0,1.0,1.0,1.0
python - <<'PY'
import json
for line in open("gpurun_out/ab_round2.jsonl"):
    j = json.loads(line)
    r = j["roofline"]
    print("%-12s %8.1f steps/s  pair %.3f ms  list %.3f ms/step  build twopass %.2f ms cell %.2f ms (%s)" % (
        j["variant"], j["value"], r["kernel_ms"], r["per_kernel_ms_per_step"]["list"], r["list_build"]["twopass_ms"], r["list_build"]["cell_ms"], r["list_build"]["in_use"]))
PY
