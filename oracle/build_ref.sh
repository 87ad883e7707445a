#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never imported by the product path.
#
# Builds the reference's own CUDA Add-RMSNorm extension (the one reference kernel
# that compiles AND imports) straight from the sources
# where they lie under /root/reference, for sm_100a, into oracle/_ref/.
#   source : /root/reference/Tools/rmsnorm/rmsnorm.cu  (+ rmsnorm.cuh)
#   output : oracle/_ref/rmsnorm_ref*.so   (python module name `rmsnorm_ref`)
# The reference's swiglu_fused extension is NOT built: it links but cannot be
# imported (swiglu_backward_cuda is declared in Tools/swiglu/swiglu.cuh:18 and
# bound in swiglu_binding.cpp:15 but never defined).
# No reference source is copied into this repository; only the binary lands in
# oracle/_ref/ (git-ignored, not gpurun-ignored).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/Tools/rmsnorm/rmsnorm.cu" ]; then
  echo "reference tree not present at $REF; keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
PY=${PYTHON:-python}
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
TARGET="$OUT/rmsnorm_ref$EXT"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/Tools/rmsnorm/rmsnorm.cu" ] && [ "$TARGET" -nt "$REF/Tools/rmsnorm/rmsnorm.cuh" ]; then
  echo "up to date: $TARGET"; exit 0
fi
INC=$($PY - <<'PY'
import sysconfig, torch.utils.cpp_extension as c
print(" ".join("-I" + p for p in c.include_paths() + [sysconfig.get_paths()["include"]]))
PY
)
TLIB=$($PY -c "import torch, os; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))")
# same flags as the reference's setup.py:14-17 (-O3 --use_fast_math -std=c++17), arch = sm_100a
nvcc -O3 --use_fast_math -std=c++17 -gencode arch=compute_100a,code=sm_100a \
  -shared -Xcompiler -fPIC -DTORCH_EXTENSION_NAME=rmsnorm_ref -DTORCH_API_INCLUDE_EXTENSION_H \
  -D_GLIBCXX_USE_CXX11_ABI=1 $INC -I"$REF/Tools/rmsnorm" \
  "$REF/Tools/rmsnorm/rmsnorm.cu" -o "$TARGET" \
  -L"$TLIB" -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda -Xlinker -rpath -Xlinker "$TLIB"
echo "built $TARGET"
