#!/bin/sh
# A/B of library variants on ONE box: tools/ab_bench.sh NAME... (NAME = default or a variants/libstar_b200.NAME.so), two rounds
# interleaved, C2 render + training step; prints rays/s, ms, SM clock.
cd "$(dirname "$0")/.."
for round in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then L=""; else L="$PWD/variants/libstar_b200.$v.so"; fi
    STAR_B200_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > /tmp/ab_$v.json 2>/tmp/ab_$v.err || { echo "$v FAILED"; tail -n 3 /tmp/ab_$v.err; continue; }
    python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ab_%s.json" % sys.argv[1]))
print("%-12s render %8.0f rays/s %7.2f ms @%4.0f MHz | train %6.3f ms  fwd %5.3f bwd %5.3f" % (sys.argv[1], d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"],
      d["train"]["ms_per_step"], d["train"]["roofline"]["calls_ms_per_step"]["mlp_forward_stash"], d["train"]["roofline"]["calls_ms_per_step"]["mlp_backward"]))
PY
  done
done
