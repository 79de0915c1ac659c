#!/bin/bash
# Builds the UNMODIFIED reference hot path from the sources where they lie under $REF (read-only) into oracle/_ref/:
#   libmaniskill_mpm.so      reference CUDA library, compiled for sm_100a (mpm/types.py:13 build line + an -arch flag)
#   libmaniskill_mpm_cpu.so  the same translation unit compiled for the host with oracle/shim/ standing in for the
#                            CUDA runtime (kernels run as host loops; used to pin oracle/mpm_oracle.c without a GPU)
# oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.  No reference source is copied into the repo.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF/mpm/csrc/integrator.cu" ]; then
  echo "build_ref.sh: $REF not present (GPU box) -- using prebuilt files in $OUT" >&2
  exit 0
fi
g++ -O2 -fPIC -shared -fopenmp -ffp-contract=off -w -x c++ -I "$HERE/shim" -I "$REF/mpm/csrc" \
    "$HERE/ref_cpu.cpp" -o "$OUT/libmaniskill_mpm_cpu.so"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin=g++ --compiler-options -fPIC -w -shared \
    "$REF/mpm/csrc/integrator.cu" -o "$OUT/libmaniskill_mpm.so"
echo "built: $(ls "$OUT")"
