#!/bin/bash
# A/B of two builds of the library on one config.  usage: tools/gpu_ab_lib.sh CFG POINTS libA.so libB.so ...
CFG=$1; PTS=$2; shift 2
for L in "$@"; do
  echo "== $L"
  SWGPU_LIB=$PWD/schwarzwald_b200/$L python bench.py --config $CFG --points $PTS --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu-baseline --no-payload 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'])"
done
