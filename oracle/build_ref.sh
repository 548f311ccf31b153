#!/usr/bin/env bash
# Build the reference's OWN src/tsdf.cu for the CPU (sequential CUDA emulation) into oracle/_ref/.
#
# Test scaffolding only. Recipe = SURVEY.md Appendix E: a scratch copy of the reference
# (under a temp dir, never committed, deleted afterwards) gets mechanical edits:
#   a. the six live  K<<<G,B>>>(args)  launch sites become  EMU_LAUNCH(K,G,B)(args)
#   b. cutil_math.h's host block (#ifndef __CUDACC__) is disabled (clashes with <math.h>)
#   c. "public:" after "class GpuTsdfGenerator {" so the driver can read h_chunks
#   d. VOXEL_PER_BLOCK is set per variant (the reference's macro, tsdf.cuh:40)
#   e. one hook call before streamOutGPU2CPU() (tsdf.cu:1585) exposing the frame's key_heap
# Output: oracle/_ref/libref_emu_vpb<V>.so  (git-ignored, travels with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VH_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
VPBS="${*:-5 8}"
[ -f "$REF/src/tsdf.cu" ] || { echo "reference not present at $REF: keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
for V in $VPBS; do
  TARGET="$OUT/libref_emu_vpb$V.so"
  if [ -f "$TARGET" ] && [ "$TARGET" -nt "$HERE/ref_emu/ref_driver.cpp" ] && [ "$TARGET" -nt "$HERE/ref_emu/emu_shim.h" ] \
     && [ "$TARGET" -nt "$HERE/build_ref.sh" ]; then
    echo "up to date: $TARGET"; continue
  fi
  TMP="$(mktemp -d /tmp/vh_ref_emu.XXXXXX)"
  trap 'rm -rf "$TMP"' EXIT
  cp -r "$REF/include" "$TMP/include"; mkdir -p "$TMP/src"; cp "$REF/src/tsdf.cu" "$TMP/src/tsdf.cu"
  chmod -R u+w "$TMP"
  # a. launch sites (both "<<<" and "<< <" spellings); the launches are single statements
  python3 - "$TMP/src/tsdf.cu" <<'PY'
import re, sys
p = sys.argv[1]; s = open(p).read()
pat = re.compile(r'^(\s*)([A-Za-z_]\w*)\s*<<\s*<\s*([^>]+?)\s*>>\s*>', re.M)
s, n = pat.subn(lambda m: f"{m.group(1)}EMU_LAUNCH({m.group(2)}, {m.group(3).split(',')[0].strip()}, {m.group(3).split(',')[1].strip()})", s)
assert n == 6, f"expected 6 live launch sites, rewrote {n}"
# e. hook
needle = "        streamOutGPU2CPU();"
assert s.count(needle) == 1
s = s.replace(needle, "        { extern void emu_hook_visible(const int3*, int); emu_hook_visible(dev_blockmap_->key_heap, *(dev_blockmap_->heap_counter)); }\n" + needle)
open(p, "w").write(s)
PY
  # b. cutil_math host block off
  sed -i '0,/^#ifndef __CUDACC__/s//#if 0/' "$TMP/include/cutil_math.h"
  # c. public members, d. VPB
  sed -i 's/^\(\s*class GpuTsdfGenerator {\)/\1 public:/' "$TMP/include/tsdf.cuh"
  sed -i "s/^#define VOXEL_PER_BLOCK .*/#define VOXEL_PER_BLOCK $V/" "$TMP/include/tsdf.cuh"
  g++ -std=c++17 -O2 -w -fPIC -shared -fopenmp -ffp-contract=off \
      -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_OMP \
      -include "$HERE/ref_emu/emu_shim.h" -I"$HERE/ref_emu" -I"$TMP/include" -I/usr/local/cuda/include \
      -x c++ "$TMP/src/tsdf.cu" -x c++ "$HERE/ref_emu/ref_driver.cpp" "$HERE/ref_emu/emu_rt.cpp" \
      -o "$TARGET"
  rm -rf "$TMP"; trap - EXIT
  echo "built $TARGET"
done
