#!/bin/bash
# A/B of experiment builds on one box:  tools/ab.sh <libdir-suffix>...   (build/x_<suffix>/; "base" = the product build)
for v in "$@"; do
  if [ "$v" = base ]; then unset PX_LIBDIR; else export PX_LIBDIR=$PWD/build/x_$v; fi
  python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'P'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json"))
    k = d["kernels"]
    print(v, "value %.4g" % d["value"], "sweep %.4f solve %.4f finish %.4f" % (k["sweep_ms_per_launch"], k["solve_ms_per_launch"], k["finish_ms_per_launch"]), "overflow", d["config"].get("worlds_overflowed"))
except Exception as e:
    print(v, "failed", e)
P
done
