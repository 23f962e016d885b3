#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference kernels for sm_100a into oracle/_ref/.
#
# The reference (nunchaku-tech/nunchaku) refuses sm_100 in its own build system (setup.py:54 allow-list),
# but the kernel translation units themselves compile for sm_100a: the INT4 path
# lowers `mma.sync.m16n8k64.s4` to emulated IMMA.16832, the NVFP4 path compiles to a trap
# (gemm_w4a4.cuh:28-32 gates it on __CUDA_ARCH__ >= 1200).  This script does NOT run the reference's
# build system and does NOT copy any reference source: it compiles the sources where they lie under
# $REF (default /root/reference) with the flags of setup.py:113-137 (arch replaced by sm_100a) and links
# them with our own shim (ref_shim.cpp: extern "C" entry points over the reference's `Tensor` type) into
#     oracle/_ref/libnunchaku_ref.so
# Outputs only under oracle/_ref/ (git-ignored, travels to the GPU box).
#
# usage: oracle/ref_build/build_ref.sh [-j N]     (≈5 min per launch TU per core; 8 cores ≈ 12 min)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REF:-/root/reference}"
OUT="$ROOT/oracle/_ref"
OBJ="$OUT/obj"
JOBS=8
while getopts "j:" o; do case $o in j) JOBS=$OPTARG;; esac; done
if [ ! -d "$REF/src/kernels/zgemm" ]; then
  echo "reference tree not present at $REF (expected on the GPU box): keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OBJ"

INC=(-I "$REF/src" -I "$REF/third_party/cutlass/include" -I "$REF/third_party/json/include"
     -I "$REF/third_party/mio/include" -I "$REF/third_party/spdlog/include")
DEFS=(-DENABLE_BF16=1 -DBUILD_NUNCHAKU=1 -UNDEBUG)
NVCC_FLAGS=("${DEFS[@]}" -std=c++20 -Xcudafe --diag_suppress=20208
  -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__
  -U__CUDA_NO_HALF2_CONVERSIONS__ -U__CUDA_NO_BFLOAT16_OPERATORS__ -U__CUDA_NO_BFLOAT16_CONVERSIONS__
  -U__CUDA_NO_BFLOAT162_OPERATORS__ -U__CUDA_NO_BFLOAT162_CONVERSIONS__
  --expt-relaxed-constexpr --expt-extended-lambda --ptxas-options=--allow-expensive-optimizations=true
  --generate-line-info -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -w)
GCC_FLAGS=("${DEFS[@]}" -std=c++20 -O1 -fPIC -w -I /usr/local/cuda/include)

# reference translation units on (or beside) the hot path, compiled where they lie
CU_SRCS=(
  src/kernels/zgemm/gemm_w4a4_launch_bf16_int4.cu
  src/kernels/zgemm/gemm_w4a4_launch_fp16_int4.cu
  src/kernels/zgemm/gemm_w4a4_launch_bf16_fp4.cu
  src/kernels/zgemm/gemm_w4a4_launch_fp16_fp4.cu
  src/kernels/zgemm/gemm_w4a4_launch_fp16_int4_fasteri2f.cu
  src/kernels/zgemm/gemm_w4a4.cu
  src/kernels/zgemm/gemm_w4a4_test.cu
  src/kernels/zgemm/attention.cu
  src/kernels/awq/gemv_awq.cu
  src/kernels/activation_kernels.cu
  src/kernels/layernorm_kernels.cu
  src/kernels/misc_kernels.cu
)
CPP_SRCS=(
  src/Linear.cpp
  src/activation.cpp
  src/layernorm.cpp
  src/Module.cpp
)

build_one() {
  local src="$1" kind="$2"
  local o="$OBJ/$(echo "$src" | tr '/' '_' | sed 's/\.[a-z]*$//').o"
  if [ -f "$o" ] && [ "$o" -nt "$REF/$src" ]; then return 0; fi
  local t0=$SECONDS
  if [ "$kind" = cu ]; then
    nvcc "${NVCC_FLAGS[@]}" "${INC[@]}" -c "$REF/$src" -o "$o.tmp" > "$o.log" 2>&1
  else
    g++ "${GCC_FLAGS[@]}" "${INC[@]}" -c "$REF/$src" -o "$o.tmp" > "$o.log" 2>&1
  fi
  mv "$o.tmp" "$o"
  echo "built $src in $((SECONDS - t0)) s"
}
export -f build_one
export OBJ REF
# arrays do not export: serialise them
export NVCC_FLAGS_S="${NVCC_FLAGS[*]}" GCC_FLAGS_S="${GCC_FLAGS[*]}" INC_S="${INC[*]}"
run() {  # run <kind> <src>: wrapper that restores the arrays in the sub-shell
  NVCC_FLAGS=($NVCC_FLAGS_S); GCC_FLAGS=($GCC_FLAGS_S); INC=($INC_S)
  build_one "$2" "$1"
}
export -f run
{
  for s in "${CU_SRCS[@]}"; do echo "cu $s"; done
  for s in "${CPP_SRCS[@]}"; do echo "cpp $s"; done
} | xargs -P "$JOBS" -L 1 bash -c 'run "$0" "$1"'

# ---- oracle/_ref/libnunchaku_ref.so: reference kernels + reference host modules + our shim -------------------------
nvcc "${NVCC_FLAGS[@]}" "${INC[@]}" -c "$HERE/ref_shim.cu" -o "$OBJ/ref_shim.o"
g++ -shared -o "$OUT/libnunchaku_ref.so" "$OBJ"/src_*.o "$OBJ/ref_shim.o" -L/usr/local/cuda/lib64 -lcudart -lcublas -Wl,-rpath,/usr/local/cuda/lib64
echo "linked $OUT/libnunchaku_ref.so"

# ---- oracle/_ref/libnunchaku_seam.so: the reference's UNMODIFIED host layer (src/Linear.cpp, Module.cpp, activation.cpp,
# layernorm.cpp objects from above) on top of OUR definitions of the zgemm.h / misc_kernels.h / activation / layernorm
# kernel entry points (nunchaku_b200/csrc/seam/*.cpp -> libnunchaku_b200.so), gemv_awq included (row N2: awq_b200.cpp).
SEAM="$ROOT/nunchaku_b200/csrc/seam"
for f in zgemm_b200 glue_b200 awq_b200; do
  g++ "${GCC_FLAGS[@]}" "${INC[@]}" -I "$ROOT/include" -c "$SEAM/$f.cpp" -o "$OBJ/seam_$f.o"
done
nvcc "${NVCC_FLAGS[@]}" "${INC[@]}" -DNREF_SEAM_BUILD=1 -c "$HERE/ref_shim.cu" -o "$OBJ/seam_shim.o"
g++ -shared -o "$OUT/libnunchaku_seam.so" "$OBJ/src_Linear.o" "$OBJ/src_Module.o" "$OBJ/src_activation.o" "$OBJ/src_layernorm.o" \
    "$OBJ/seam_awq_b200.o" "$OBJ/seam_zgemm_b200.o" "$OBJ/seam_glue_b200.o" "$OBJ/seam_shim.o" \
    -L"$ROOT/nunchaku_b200/_lib" -lnunchaku_b200 -L/usr/local/cuda/lib64 -lcudart -lcublas \
    -Wl,-rpath,/usr/local/cuda/lib64 -Wl,-rpath,'$ORIGIN/../../nunchaku_b200/_lib' -Wl,--no-undefined
echo "linked $OUT/libnunchaku_seam.so"

# ---- oracle/_ref/pyseam/_C.so: the `nunchaku._C.ops` pybind surface (reference nunchaku/csrc/ops.h + src/interop/torch.cpp, compiled
# where they lie) on top of the same seam objects.  Needs the torch headers of this interpreter.
PY="${PYTHON:-python}"
TORCH_INC=$($PY - <<'PYEOF'
from torch.utils.cpp_extension import include_paths
import sysconfig
print(" ".join("-I" + p for p in include_paths() + [sysconfig.get_paths()["include"]]))
PYEOF
)
TORCH_LIB=$($PY -c "import torch, os; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))")
PYB_INC=$($PY -c "import pybind11; print(pybind11.get_include())" 2>/dev/null || true)
mkdir -p "$OUT/pyseam"
TFLAGS=("${DEFS[@]}" -std=c++20 -O1 -fPIC -w -I /usr/local/cuda/include $TORCH_INC ${PYB_INC:+-I $PYB_INC} -DTORCH_EXTENSION_NAME=_C -D_GLIBCXX_USE_CXX11_ABI=1)
if [ ! -f "$OBJ/py_torch_interop.o" ] || [ "$REF/src/interop/torch.cpp" -nt "$OBJ/py_torch_interop.o" ]; then
  g++ "${TFLAGS[@]}" "${INC[@]}" -c "$REF/src/interop/torch.cpp" -o "$OBJ/py_torch_interop.o"
fi
g++ "${TFLAGS[@]}" "${INC[@]}" -I "$REF/nunchaku/csrc" -I "$ROOT/include" -c "$SEAM/pybind_ops.cpp" -o "$OBJ/py_pybind_ops.o"
g++ -shared -o "$OUT/pyseam/_C.so" "$OBJ/py_pybind_ops.o" "$OBJ/py_torch_interop.o" "$OBJ/seam_zgemm_b200.o" "$OBJ/seam_awq_b200.o" \
    -L"$ROOT/nunchaku_b200/_lib" -lnunchaku_b200 -L"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python \
    -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,/usr/local/cuda/lib64 -Wl,-rpath,"$TORCH_LIB" -Wl,-rpath,'$ORIGIN/../../../nunchaku_b200/_lib'
echo "linked $OUT/pyseam/_C.so"
