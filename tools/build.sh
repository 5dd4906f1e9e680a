#!/bin/sh
# (re)builds libstar_b200.so in-tree; extra nvcc flags through STAR_B200_NVCC_EXTRA (e.g. -DSTAR_TC_DEBUG for the
# bottleneck-experiment switches of csrc/mlp_tc.cu)
cd "$(dirname "$0")/.."
python - "$@" <<'PY'
import importlib.util, sys
spec = importlib.util.spec_from_file_location("_star_build", "3d-mot-using-neural-radiance-fields_b200/_build.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
print(m.build(force="--force" in sys.argv, verbose="-v" in sys.argv))
PY
