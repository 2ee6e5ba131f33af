#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds the UNMODIFIED reference rasterizer for sm_100a from
# the sources where they lie under /root/reference into oracle/_ref/ (git-ignored,
# travels to the GPU box with gpurun).  Only runs where /root/reference exists.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GS2MESH_REFERENCE:-/root/reference}"
DGR="$REF/third_party/gaussian-splatting/submodules/diff-gaussian-rasterization"
OUT="$HERE/_ref"
if [ ! -d "$DGR/cuda_rasterizer" ]; then
  echo "build_ref: $DGR not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
# -include cstdint: rasterizer_impl.h uses std::uintptr_t / uint32_t without it (gcc 13).
NVFLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -include cstdint
         -I "$DGR" -I "$DGR/third_party/glm" -Xcompiler -fPIC -w)
objs=()
for f in forward rasterizer_impl backward; do
  if [ ! -f "$OUT/$f.o" ] || [ "$DGR/cuda_rasterizer/$f.cu" -nt "$OUT/$f.o" ]; then
    nvcc "${NVFLAGS[@]}" -c "$DGR/cuda_rasterizer/$f.cu" -o "$OUT/$f.o" &
  fi
  objs+=("$OUT/$f.o")
done
nvcc "${NVFLAGS[@]}" -c "$HERE/ref_dgr_capi.cu" -o "$OUT/ref_dgr_capi.o" &
wait
nvcc -shared -o "$OUT/libref_dgr.so" "${objs[@]}" "$OUT/ref_dgr_capi.o" -lcudart
echo "build_ref: wrote $OUT/libref_dgr.so"
