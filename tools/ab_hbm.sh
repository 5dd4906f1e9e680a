#!/bin/sh
# A/B of library variants on the HBM microbenchmark (ONE box): tools/ab_hbm.sh FILTER NAME...  (NAME = default or a
# variants/libstar_b200.NAME.so); prints the lines of tools/hbm_microbench.py that match FILTER, two rounds interleaved.
cd "$(dirname "$0")/.."
F=$1; shift
for round in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then L=""; else L="$PWD/variants/libstar_b200.$v.so"; fi
    echo "== $v"; STAR_B200_LIB=$L ONLY="$F" timeout 300 python tools/hbm_microbench.py 2>&1 | grep -E "$F"
  done
done
