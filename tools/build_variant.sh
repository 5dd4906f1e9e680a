#!/bin/sh
# tools/build_variant.sh NAME "<extra nvcc flags>": builds the library with extra flags into variants/libstar_b200.NAME.so
# (travels with gpurun; select with STAR_B200_LIB=variants/libstar_b200.NAME.so) and restores the default build.
set -e
cd "$(dirname "$0")/.."
P=3d-mot-using-neural-radiance-fields_b200
mkdir -p variants
STAR_B200_NVCC_EXTRA="$2" python -c "import importlib.util,sys; s=importlib.util.spec_from_file_location('b','$P/_build.py'); m=importlib.util.module_from_spec(s); s.loader.exec_module(m); m.build(force=True)"
cp $P/libstar_b200.so variants/libstar_b200.$1.so
python -c "import importlib.util,sys; s=importlib.util.spec_from_file_location('b','$P/_build.py'); m=importlib.util.module_from_spec(s); s.loader.exec_module(m); m.build(force=True)"
echo variants/libstar_b200.$1.so
