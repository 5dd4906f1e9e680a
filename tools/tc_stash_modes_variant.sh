cd /root/repo
L=$PWD/variants/libstar_b200.dbg.so
for m in 0 8 16 32 40 56; do
  echo "mode $m"; STAR_B200_LIB=$L STAR_TC_DEBUG_CYCLES=1 STAR_TC_DEBUG_MODE=$m timeout 200 python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep "star_tc" | sort | uniq -c | sort -rn | head -4
done
echo "render"; STAR_B200_LIB=$L STAR_TC_DEBUG_CYCLES=1 timeout 200 python bench.py --hw 200 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-train-extra 2>&1 | grep "star_tc" | sort | uniq -c | sort -rn | head -4
