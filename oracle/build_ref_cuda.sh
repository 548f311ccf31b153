#!/usr/bin/env bash
# Build the reference's OWN src/tsdf.cu as a real CUDA library for sm_100a into oracle/_ref/ (baseline measurement:
# "the reference's own CUDA build on the same box"). Test/bench scaffolding only; needs /root/reference
# (run here, the .so travels to the GPU box with gpurun). Edits to the scratch copy, all by script:
#   * "public:" after "class GpuTsdfGenerator {" so the driver can read h_chunks / counters (introspection only)
#   * VOXEL_PER_BLOCK per variant (reference macro, tsdf.cuh:40); variant "8c1" additionally sets the two hash-table
#     shapes at tsdf.cu:1488,1492 to 1048576 buckets x 4 entries with a 1,000,000-block heap (BASELINE config 1 streams
#     ~453k blocks per frame, more than the reference's 400,000-block heap)
# Flags = the reference's own (-O3 -use_fast_math, CMakeLists.txt:13) + -arch=sm_100a; the 8^3 variants need
# -maxrregcount=128 because the kernels are launched with 512 threads (tsdf.cu:1527) and would otherwise use 255 registers.
# -DNDEBUG: on a B200 the reference's device-side assert in its block allocator (impl/blockalloc.h:36) fires in the very first
# frame (a reader sees an entry whose block_index is not yet published: vhashing.h:408-428 writes the key first); with
# asserts compiled out that read returns slot 0 (reserved, valid memory) and the run continues, as it would have on the
# asserts-off builds the reference's CMake Release flags produce.
# GL is stubbed (oracle/ref_emu/GL), __cudaSafeCall is replaced by a glog-free one (safecall_stub.cpp).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VH_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
[ -f "$REF/src/tsdf.cu" ] || { echo "reference not present at $REF: keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
for VAR in ${*:-8 8c1}; do
  V="${VAR%%c1}"
  TARGET="$OUT/libref_cuda_vpb$VAR.so"
  if [ -f "$TARGET" ] && [ "$TARGET" -nt "$HERE/ref_emu/ref_driver.cpp" ] && [ "$TARGET" -nt "$HERE/build_ref_cuda.sh" ]; then echo "up to date: $TARGET"; continue; fi
  TMP="$(mktemp -d /tmp/vh_ref_cuda.XXXXXX)"
  cp -r "$REF/include" "$TMP/include"; mkdir -p "$TMP/src"; cp "$REF/src/tsdf.cu" "$TMP/src/tsdf.cu"; chmod -R u+w "$TMP"
  sed -i 's/^\(\s*class GpuTsdfGenerator {\)/\1 public:/' "$TMP/include/tsdf.cuh"
  sed -i "s/^#define VOXEL_PER_BLOCK .*/#define VOXEL_PER_BLOCK $V/" "$TMP/include/tsdf.cuh"
  if [ "$VAR" != "$V" ]; then
    python3 - "$TMP/src/tsdf.cu" <<'PY'
import re, sys
p = sys.argv[1]; s = open(p).read()
s, n = re.subn(r"(?m)^(\s*dev_blockmap\w*\s*=\s*new [^\n]*?)\(\s*100001\s*,\s*4\s*,\s*400000\s*,", r"\1(1048576, 4, 1000000,", s)
assert n == 2, f"expected the two table constructions, rewrote {n}"
open(p, "w").write(s)
PY
  fi
  cp "$HERE/ref_emu/ref_driver.cpp" "$TMP/ref_driver.cu"
  cat > "$TMP/safecall_stub.cu" <<'CPP'
#include <cuda_runtime.h>
#include <cstdio>
void __cudaSafeCall(cudaError err, const char* file, const int line) {
  if (err != cudaSuccess) { fprintf(stderr, "cudaSafeCall failed at %s:%d: %s\n", file, line, cudaGetErrorString(err)); throw "CUDA Error"; }
}
CPP
  MAXR=""; [ "$V" = "8" ] && MAXR="-maxrregcount=128"
  nvcc -std=c++17 -O3 -use_fast_math -DNDEBUG -gencode arch=compute_100a,code=sm_100a $MAXR -w -Xcompiler -fPIC -shared \
       -I"$HERE/ref_emu" -I"$TMP/include" "$TMP/src/tsdf.cu" "$TMP/ref_driver.cu" "$TMP/safecall_stub.cu" -o "$TARGET"
  rm -rf "$TMP"
  echo "built $TARGET"
done
