# needs the -DSTAR_TC_DEBUG build (the switches are compiled out of the production library); restores it afterwards
STAR_B200_NVCC_EXTRA=-DSTAR_TC_DEBUG tools/build.sh --force > /dev/null
for m in 0 1 2 4 3 6; do
  echo "mode $m"; STAR_TC_DEBUG_CYCLES=1 STAR_TC_DEBUG_MODE=$m timeout 200 python bench.py --hw 200 --steps 1 --warmup 1 --no-cpu-baseline --no-train-extra 2>&1 | grep "star_tc" | sort | uniq -c | sort -rn | head -4
done
tools/build.sh --force > /dev/null
