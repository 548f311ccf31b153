/* CUDA-on-CPU emulation shim (test scaffolding, NOT product code).
 *
 * Force-included (-include) in front of a scratch copy of the reference's src/tsdf.cu so
 * that plain g++ can compile it and run every __global__ kernel sequentially, one
 * thread at a time in grid order. No reference kernel uses shared memory or
 * __syncthreads, so sequential execution is the reference's exact single-thread meaning
 * with IEEE host floats. Recipe: SURVEY.md Appendix E.
 */
#pragma once
#include <cuda_runtime.h>
#include <device_launch_parameters.h>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cassert>
#include <algorithm>

struct emu_idx3 { unsigned int x, y, z; };
extern emu_idx3 emu_threadIdx, emu_blockIdx, emu_blockDim, emu_gridDim;
#define threadIdx emu_threadIdx
#define blockIdx  emu_blockIdx
#define blockDim  emu_blockDim
#define gridDim   emu_gridDim

static inline int atomicCAS(int* a, int cmp, int v) { int o = *a; if (o == cmp) *a = v; return o; }
static inline int atomicExch(int* a, int v) { int o = *a; *a = v; return o; }
static inline int atomicAdd(int* a, int v) { int o = *a; *a = o + v; return o; }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { unsigned o = *a; *a = o + v; return o; }
static inline int atomicSub(int* a, int v) { int o = *a; *a = o - v; return o; }
static inline int atomicMax(int* a, int v) { int o = *a; if (v > o) *a = v; return o; }
static inline void __threadfence() {}
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }

/* The reference's kernels call unqualified min/max on floats (tsdf.cu:2123-2124) and ints;
 * cutil_math.h's host block (int-only min/max) is disabled in the scratch copy. */
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

/* The call operator takes exactly the kernel's parameter types, so arguments convert at the
 * call site as they would for a real launch (e.g. HashTable -> HashTableBase by-value slicing). */
template <class... P>
struct EmuLaunch {
  void (*k)(P...); dim3 g, b;
  void operator()(P... args) const {
    emu_gridDim = {g.x, g.y, g.z};
    emu_blockDim = {b.x, b.y, b.z};
    for (unsigned bz = 0; bz < g.z; bz++) for (unsigned by = 0; by < g.y; by++) for (unsigned bx = 0; bx < g.x; bx++)
      for (unsigned tz = 0; tz < b.z; tz++) for (unsigned ty = 0; ty < b.y; ty++) for (unsigned tx = 0; tx < b.x; tx++) {
        emu_blockIdx = {bx, by, bz};
        emu_threadIdx = {tx, ty, tz};
        k(args...);
      }
  }
};
template <class... P> static inline EmuLaunch<P...> emu_make_launch(void (*k)(P...), dim3 g, dim3 b) { return EmuLaunch<P...>{k, g, b}; }
#define EMU_LAUNCH(k, g, b) emu_make_launch(k, dim3(g), dim3(b))
