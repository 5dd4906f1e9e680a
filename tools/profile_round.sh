#!/bin/sh
# One gpurun call: the round's ncu evidence for the kernels as they are NOW.  tools/profile_round.sh TAG
#   1. the bench command plain (must exit 0), then its launch list (cold-cache, serialised: compare SHARES);
#   2. `--set full` of the two inference launches of one render step (coarse + fine MLP);
#   3. `--set full` of the tensor-core training kernels of one 4096-ray step.
# Outputs land in gpurun_out/TAG_*; copy the summaries into profiles/.
TAG=${1:-r2i}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --hw 400 --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
TRAIN="python bench.py --mode train --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -n 5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# warm-up = 3 render steps of 2 MLP launches each
ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_tc -s 6 -c 2 -f -o gpurun_out/${TAG}_fwd $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i gpurun_out/${TAG}_fwd.ncu-rep --page raw --csv > gpurun_out/${TAG}_fwd_raw.csv 2>/dev/null
$TRAIN > gpurun_out/${TAG}_train_plain.json 2> gpurun_out/${TAG}_train_plain.err || { echo "plain train run failed"; tail -n 5 gpurun_out/${TAG}_train_plain.err; exit 1; }
# warm-up = 3 steps of (2 stash forwards + 2 dX + 2 dW)
ncu --set full --clock-control none -k regex:"mlp_fwd_tc|mlp_bwd_tc|dw_tc_kernel" -s 18 -c 6 -f -o gpurun_out/${TAG}_train $TRAIN > gpurun_out/${TAG}_ncu3.log 2>&1
ncu -i gpurun_out/${TAG}_train.ncu-rep --page raw --csv > gpurun_out/${TAG}_train_raw.csv 2>/dev/null
ls -la gpurun_out | tail -n 15
cat gpurun_out/${TAG}_plain.json | head -c 600
