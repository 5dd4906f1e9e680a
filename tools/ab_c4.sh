#!/bin/sh
# A/B of library variants on the C4 training step: tools/ab_c4.sh NAME... (default or variants/libstar_b200.NAME.so), two rounds
cd "$(dirname "$0")/.."
for round in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then L=""; else L="$PWD/variants/libstar_b200.$v.so"; fi
    STAR_B200_LIB=$L timeout 600 python bench.py --workload c4 --mode train --steps 5 --warmup 3 --no-extras --no-cpu-baseline > /tmp/abc4_$v.json 2>/tmp/abc4_$v.err || { echo "$v FAILED"; tail -n 3 /tmp/abc4_$v.err; continue; }
    python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/abc4_%s.json" % sys.argv[1]))
print("%-12s c4 train %8.0f rays/s %7.2f ms @%4.0f MHz | %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["calls_ms_per_step"]))
PY
  done
done
