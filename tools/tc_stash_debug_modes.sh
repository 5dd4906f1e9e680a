# needs the -DSTAR_TC_DEBUG build (the switches are compiled out of the production library); restores it afterwards
STAR_B200_NVCC_EXTRA=-DSTAR_TC_DEBUG tools/build.sh --force > /dev/null
# Where does the training (stash) variant of the forward kernel lose time against the inference variant?
# Debug bits (results are garbage, timing only): 8 = no stash bulk stores, 16 = no ReLU bit masks, 32 = no stash_done waits.
for m in 0 8 16 32 56; do
  echo "mode $m"; STAR_TC_DEBUG_CYCLES=1 STAR_TC_DEBUG_MODE=$m timeout 200 python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep "star_tc" | sort | uniq -c | sort -rn | head -3
done
tools/build.sh --force > /dev/null
