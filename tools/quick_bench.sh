#!/bin/bash
# usage: tools/quick_bench.sh TAG [ENV=VAL ...] -- [bench args]; prints one summary line (GPU box helper)
tag=$1; shift
envs=()
while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done
shift
env "${envs[@]}" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs "$@" > gpurun_out/qb_$tag.json 2> gpurun_out/qb_$tag.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/qb_{tag}.json"))
    s = d["solver_stats"]
    print(tag, "value %.3fM" % (d["value"] / 1e6), "mpc_ms %.2f" % s["mpc_kernel_ms_per_tick"], "nfac %.2f" % s["factorisations_per_tick"], s["hot_path"])
except Exception as e:
    print(tag, "FAILED", e)
    print(open(f"gpurun_out/qb_{tag}.err").read()[-800:])
PY
