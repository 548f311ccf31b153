// tests/emu/cuda_runtime.h — a host stand-in for the CUDA device model. TEST INFRASTRUCTURE ONLY.
//
// Lets g++ compile the engine's kernel sources (voxel-hashing-sdf_b200/csrc/*.cu, included as text with -DVH_HOST_EMU)
// and execute them on the CPU, lane by lane, so that the `-m "not gpu"` tests can check the very code the GPU runs
// against the oracle. It is found as <cuda_runtime.h> through -I tests/emu; nothing in the product includes or links it.
//
// Execution model: a CTA is a set of fibres (ucontext), one per CUDA thread, run round-robin by a scheduler on ONE OS
// thread; CTAs run one after the other. A fibre runs until it reaches a collective (__shfl*_sync, __ballot_sync,
// __any_sync, __match_any_sync, __syncwarp, __syncthreads) and yields there until every member of the warp / CTA has
// arrived. That is the converged-warp subset of the SIMT model, which is all these kernels use (full masks). Atomics
// are plain read-modify-writes (nothing runs concurrently). Floating point: the intrinsics map onto IEEE binary32
// operations of the host (-ffp-contract=off, fmaf); MUFU.RCP is emulated by the correctly rounded reciprocal perturbed
// by up to +-1 ulp (seeded), which exercises the guard bands the kernels put around approximate results.
#pragma once
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __CUDACC__ 1
#define VH_HOST_EMU 1

// ---- vector types ---------------------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct int3 { int x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct uchar3 { unsigned char x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int3 make_int3(int x, int y, int z) { return int3{x, y, z}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;

// ---- fibres ---------------------------------------------------------------------------------------------------------
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>      // fibre switches must be announced to AddressSanitizer (make -C tests/emu asan)
#define EMU_ASAN 1
#endif

namespace emu {

struct Group { int size = 0, arrived = 0; unsigned gen = 0; unsigned long long slot[2][32]; unsigned parity = 0; };
struct Fibre {
  ucontext_t ctx;
  char* stack = nullptr;
  uint3 tid;
  bool done = false;
  unsigned warp_parity = 0;   // which exchange buffer this lane uses for its next warp collective
  void* asan_fake = nullptr;  // AddressSanitizer's fake-stack handle while the fibre is switched out
};
struct Cta {
  std::vector<Fibre> fibres;
  std::vector<Group> warps;
  Group all;
  ucontext_t sched;
  Fibre* cur = nullptr;
  uint3 bid;
  dim3 bdim, gdim;
  void (*entry)(void*) = nullptr;
  void* arg = nullptr;
  char* dyn_smem = nullptr;      // dynamic shared memory of the CTA (third launch parameter)
  const void* sched_stack = nullptr; size_t sched_stack_size = 0;     // the scheduler's (OS thread's) stack, for AddressSanitizer
  size_t fibre_stack_bytes = 0;
};
extern Cta* g_cta;
extern unsigned long long g_collectives, g_events;
extern unsigned g_rcp_seed;

inline void yield() {
#ifdef EMU_ASAN
  Fibre* f = g_cta->cur;
  __sanitizer_start_switch_fiber(&f->asan_fake, g_cta->sched_stack, g_cta->sched_stack_size);
  swapcontext(&f->ctx, &g_cta->sched);
  __sanitizer_finish_switch_fiber(f->asan_fake, nullptr, nullptr);
#else
  swapcontext(&g_cta->cur->ctx, &g_cta->sched);
#endif
}
inline void barrier(Group& g) {
  const unsigned my = g.gen;
  if (++g.arrived == g.size) { g.arrived = 0; g.gen++; g_events++; } else while (g.gen == my) yield();
}
inline int linear_tid() { const uint3& t = g_cta->cur->tid; return (int)(t.x + g_cta->bdim.x * (t.y + g_cta->bdim.y * t.z)); }
inline Group& my_warp() { return g_cta->warps[linear_tid() >> 5]; }
inline int my_lane() { return linear_tid() & 31; }

// every lane deposits v, waits for the warp, and then sees all 32 deposits; two buffers alternate so that a lane that
// races ahead to the next collective cannot overwrite values its siblings have not read yet
inline const unsigned long long* exchange(unsigned long long v) {
  Group& w = my_warp();
  Fibre* f = g_cta->cur;
  const unsigned p = f->warp_parity; f->warp_parity ^= 1u;
  w.slot[p][my_lane()] = v;
  g_collectives++;
  barrier(w);
  return w.slot[p];
}

void trampoline();
// runs kernel(arg) for every thread of every CTA of the grid, CTAs one after the other
void run_grid(dim3 grid, dim3 block, void (*entry)(void*), void* arg, size_t dyn_smem_bytes = 0);

}  // namespace emu

#define threadIdx (emu::g_cta->cur->tid)
#define blockIdx (emu::g_cta->bid)
#define blockDim (emu::g_cta->bdim)
#define gridDim (emu::g_cta->gdim)

// ---- collectives ----------------------------------------------------------------------------------------------------
static inline void emu_check_mask(unsigned m) { if (m != 0xffffffffu) { fprintf(stderr, "emu: partial-mask collective (0x%x) is not modelled\n", m); abort(); } }
template <typename T> static inline T emu_from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> static inline unsigned long long emu_to_bits(T v) { static_assert(sizeof(T) <= 8, "shuffle of a wide type"); unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  emu_check_mask(m);
  const int lane = emu::my_lane(), base = lane & ~(width - 1);
  return emu_from_bits<T>(emu::exchange(emu_to_bits(v))[base + (src & (width - 1))]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  emu_check_mask(m);
  const int lane = emu::my_lane();
  const unsigned long long* s = emu::exchange(emu_to_bits(v));
  const int src = lane ^ x;
  return emu_from_bits<T>(s[(src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane]);
}
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  emu_check_mask(m);
  const int lane = emu::my_lane();
  const unsigned long long* s = emu::exchange(emu_to_bits(v));
  const int src = lane + (int)d;
  return emu_from_bits<T>(s[(src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane]);
}
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  emu_check_mask(m);
  const int lane = emu::my_lane();
  const unsigned long long* s = emu::exchange(emu_to_bits(v));
  const int src = lane - (int)d;
  return emu_from_bits<T>(s[src >= (lane & ~(width - 1)) ? src : lane]);
}
static inline unsigned __ballot_sync(unsigned m, int pred) {
  emu_check_mask(m);
  const unsigned long long* s = emu::exchange(pred ? 1ull : 0ull);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= s[i] ? (1u << i) : 0u;
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
template <typename T> static inline unsigned __match_any_sync(unsigned m, T v) {
  emu_check_mask(m);
  const unsigned long long mine = emu_to_bits(v);
  const unsigned long long* s = emu::exchange(mine);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= s[i] == mine ? (1u << i) : 0u;
  return r;
}
static inline void __syncwarp(unsigned m = 0xffffffffu) { emu_check_mask(m); emu::barrier(emu::my_warp()); }
static inline void __syncthreads() { emu::barrier(emu::g_cta->all); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

// ---- atomics (nothing runs concurrently) ----------------------------------------------------------------------------
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U> static inline T atomicAdd_system(T* p, U v) { return atomicAdd(p, v); }
template <typename T, typename U> static inline T atomicSub(T* p, U v) { const T o = *p; *p = (T)(o - (T)v); return o; }
template <typename T, typename U> static inline T atomicOr(T* p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U> static inline T atomicAnd(T* p, U v) { const T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename U> static inline T atomicMax(T* p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicMin(T* p, U v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicExch(T* p, U v) { const T o = *p; *p = (T)v; return o; }
template <typename T, typename U, typename V> static inline T atomicCAS(T* p, U cmp, V v) { const T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---- loads ----------------------------------------------------------------------------------------------------------
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }

// ---- arithmetic intrinsics: single IEEE binary32 operations (the harness is built with -ffp-contract=off) ------------
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline float __fadd_rd(float a, float b) {
  const int old = fegetround(); fesetround(FE_DOWNWARD);
  volatile float x = a, y = b; volatile float r = x + y;
  fesetround(old); return r;
}
static inline float __fadd_ru(float a, float b) {
  const int old = fegetround(); fesetround(FE_UPWARD);
  volatile float x = a, y = b; volatile float r = x + y;
  fesetround(old); return r;
}
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline int __float2int_rz(float a) {   // saturating, NaN -> 0 like cvt.rzi.s32.f32
  if (a != a) return 0;
  if (a >= 2147483648.0f) return 2147483647;
  if (a <= -2147483648.0f) return -2147483647 - 1;
  return (int)a;
}
static inline int __float2int_rn(float a) { if (a != a) return 0; if (a >= 2147483648.0f) return 2147483647; if (a <= -2147483648.0f) return -2147483647 - 1; return (int)nearbyintf(a); }
static inline int __float2int_rd(float a) { return __float2int_rz(floorf(a)); }
static inline float __double2float_rn(double a) { return (float)a; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const unsigned long long src = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned sel = (s >> (4 * i)) & 0xF;
    unsigned b = (unsigned)(src >> (8 * (sel & 7))) & 0xFF;
    if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
    r |= b << (8 * i);
  }
  return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline int __mul24(int a, int b) { return a * b; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float fminf_cuda(float a, float b) { return fminf(a, b); }

// MUFU.RCP stand-in: 1/x correctly rounded, then moved by -1, 0 or +1 ulp (pseudo-random): within the 1-ulp error the
// kernels assume for rcp.approx.ftz.f32. Denormal inputs/outputs flush to zero like .ftz.
static inline float emu_rcp_approx(float x) {
  if (x != x) return x;
  if (fabsf(x) < 1.17549435e-38f) return copysignf(INFINITY, x);
  float r = 1.0f / x;
  if (!(fabsf(r) >= 1.17549435e-38f)) return copysignf(0.0f, r);
  if (isinf(r)) return r;
  emu::g_rcp_seed = emu::g_rcp_seed * 1664525u + 1013904223u;
  const unsigned k = (emu::g_rcp_seed >> 16) % 3u;
  unsigned b = __float_as_uint(r);
  if (k == 1) b += 1; else if (k == 2) b -= 1;
  return __uint_as_float(b);
}
