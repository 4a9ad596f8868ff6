#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle) -- builds oracle/_ref/libgsref.so: the UNMODIFIED reference
# rasterizer (forward.cu, backward.cu, rasterizer_impl.cu) compiled for sm_100a from the sources
# where they lie under /root/reference, behind the plain-pointer shim ref_build/ref_shim.cu.
# The reference's own build system (setup.py / CMake) is not run and no source is copied.
# Output goes only into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GSREF_SRC:-/root/reference/submodules/diff-gaussian-rasterization}"
OUT="$HERE/_ref"
if [ ! -d "$REF/cuda_rasterizer" ]; then
	echo "reference sources not present at $REF (expected on the GPU box): keeping prebuilt $OUT/libgsref.so" >&2
	exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -w
	-I"$REF" -I"$REF/third_party/glm")
# NOTE: never -I$REF/cuda_rasterizer: its math.h would shadow <math.h>.
pids=()
for tu in forward backward rasterizer_impl; do
	"$NVCC" "${FLAGS[@]}" -include "$HERE/ref_build/ref_prelude.h" \
		-c "$REF/cuda_rasterizer/$tu.cu" -o "$OUT/$tu.o" &
	pids+=($!)
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_build/ref_shim.cu" -o "$OUT/ref_shim.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libgsref.so" \
	"$OUT/forward.o" "$OUT/backward.o" "$OUT/rasterizer_impl.o" "$OUT/ref_shim.o" -lcudart
rm -f "$OUT"/*.o
echo "built $OUT/libgsref.so"
