#!/bin/sh
# Debug build of the tensor-core forward kernel with clock64 stamps (-DSTAR_TC_TIMELINE), run on a GPU box:
# prints, for the second tile of CTA 0, when the issuer saw each layer's first operand block / issued its last K-block
# and when epilogue warp 0 saw the accumulator / finished the layer.  Restores the production library afterwards.
set -e
cd "$(dirname "$0")/.."
P=3d-mot-using-neural-radiance-fields_b200
cp $P/libstar_b200.so /tmp/libstar_b200.prod.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DSTAR_TC_DEBUG -DSTAR_TC_TIMELINE ${TL_LAYER:+-DTL_LAYER=$TL_LAYER} -c $P/csrc/mlp_tc.cu -o /tmp/mlp_tc_tl.o
OBJS=$(ls $P/build/*.o | grep -v mlp_tc.o)
nvcc -shared -o $P/libstar_b200.so $OBJS /tmp/mlp_tc_tl.o -gencode arch=compute_100a,code=sm_100a -lcudart
STAR_TC_DEBUG_CYCLES=1 R=${R:-8192} python tools/tc_microbench.py 2>&1 | grep star_tc | head -20
cp /tmp/libstar_b200.prod.so $P/libstar_b200.so
