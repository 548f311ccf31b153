// vh_alloc.cu — visible-block discovery and allocation for one depth frame.
//
// Replaces three reference stages with one kernel and no host round trips:
//   streamInCPU2GPU + streamInCPU2GPUKernel  (/root/reference/src/tsdf.cu:277-457, :208-216)  -> the candidate
//       test is evaluated analytically per block (chunk cube + chunk sphere) instead of uploading chunks;
//   HashAssignKernel                         (tsdf.cu:2088-2238)  -> same rays, same 3-D DDA, same frustum test;
//   getHeapCounterKernel + 4 memcpys         (tsdf.cu:2244-2251, :2318-2337) -> a device-side counter.
//
// Shape: one CTA owns a tile of 4x2 sampled pixels (8 rays); small tiles keep every SM busy (3,072 rays in all).
//   Phase A  8 lanes march their rays through the block grid (pure ALU, sequential per ray because the
//            reference accumulates tmax by repeated float addition) and drop the visited block keys into
//            shared memory, keys[step][ray].
//   Phase B  all 8 warps sweep that list; a warp sees the 8 neighbouring rays of 4 consecutive steps, so equal keys are
//            folded with __match_any_sync, the group leader claims the entry with one 64-bit atomicCAS,
//            pool slots for new blocks are popped with one atomicSub per warp (ballot/popc ranks), and blocks
//            first seen this frame (atomicExch on the entry's frame stamp) are compacted into the visible list
//            with one atomicAdd per warp.
#include <cstdlib>

#include "vh_engine.h"
#include "vh_math.cuh"

namespace vh {

constexpr int RAYS_X = 4, RAYS_Y = 2, RAYS = RAYS_X * RAYS_Y;   // 8 rays per CTA: 384 CTAs at 640x480 / stride 10, ~3 key rounds each
constexpr int ALLOC_THREADS = 256;

// chunk membership of the reference's stream-in stage
__device__ __forceinline__ bool chunk_is_candidate(const StaticParams& S, const FrameParams& F, int x, int y, int z) {
  if (x < F.cstart[0] || x > F.cend[0] || y < F.cstart[1] || y > F.cend[1] || z < F.cstart[2] || z > F.cend[2]) return false;
  // make_float3(((float)x + 0.5) * chunk_size, ...): double arithmetic narrowed to float (tsdf.cu:168)
  const double cs = (double)S.chunk_size;
  const float ccx = __double2float_rn(dmul(dadd((double)i2f(x), 0.5), cs));
  const float ccy = __double2float_rn(dmul(dadd((double)i2f(y), 0.5), cs));
  const float ccz = __double2float_rn(dmul(dadd((double)i2f(z), 0.5), cs));
  const float vx = fsub(F.fc[0], ccx), vy = fsub(F.fc[1], ccy), vz = fsub(F.fc[2], ccz);
  const float l = fsqrt(fadd(fadd(fmul(vx, vx), fmul(vz, vz)), fmul(vy, vy)));   // x, z, y order (tsdf.cu:174)
  return l <= fabsf(F.chunk_test_radius);
}

__device__ __forceinline__ int block_to_chunk(int b, float bpc) { return __float2int_rd(fdiv(i2f(b), bpc)); }   // tsdf.cu:256-260

// isBlockInCameraFrustum on the block's minimum corner (tsdf.cu:2013-2064)
__device__ __forceinline__ bool block_in_frustum(const StaticParams& S, const FrameParams& F, int bx, int by, int bz) {
  const Float3 c = world_to_cam(F.c2w, fmul(i2f(bx), S.block_size), fmul(i2f(by), S.block_size), fmul(i2f(bz), S.block_size));
  const float u = fadd(fdiv(fmul(c.x, S.fx), c.z), S.cx);
  const float v = fadd(fdiv(fmul(c.y, S.fy), c.z), S.cy);
  const float wm1 = fsub(i2f(S.W), 1.0f), hm1 = fsub(i2f(S.H), 1.0f);
  float ix = fdiv(fsub(fmul(2.0f, u), wm1), wm1);
  float iy = fdiv(fsub(hm1, fmul(2.0f, v)), hm1);
  float iz = fdiv(fsub(c.z, S.min_depth), fsub(S.max_depth, S.min_depth));
  const float k = 0.95f;
  ix = fmul(ix, k); iy = fmul(iy, k); iz = fmul(iz, k);
  return !(ix < -1.0f || ix > 1.0f || iy < -1.0f || iy > 1.0f || iz < 0.0f || iz > 1.0f);
}

__device__ __forceinline__ float sign_f(float v) { return (float)((0.0f < v) - (v < 0.0f)); }

constexpr int CHUNK_STEPS = 25;                       // DDA steps marched per pipeline stage
constexpr int CHUNK_KEYS = CHUNK_STEPS * RAYS;        // 200 keys per stage <= 224 consumer threads
static_assert(CHUNK_KEYS <= ALLOC_THREADS - 32, "one consumer round per stage");

// per-ray DDA state, kept in registers of the marching lanes across pipeline stages
struct RayState {
  int cur[3], bound[3], istep[3];
  float tmax[3], tdel[3];
  bool alive;
};

__device__ __forceinline__ void ray_setup(const StaticParams& S, const FrameParams& F, const float* __restrict__ depth, int rx, int ry, RayState& R) {
  const unsigned x = (unsigned)rx * (unsigned)S.stride, y = (unsigned)ry * (unsigned)S.stride;
  R.alive = rx < S.nrx && ry < S.nry;
#pragma unroll
  for (int a = 0; a < 3; a++) { R.cur[a] = 0; R.bound[a] = 0; R.istep[a] = 0; R.tmax[a] = 0.0f; R.tdel[a] = 0.0f; }
  if (R.alive) {
    // depth[x*width + y] with x = column: the reference's transposed gate (tsdf.cu:2114).
    // Indices past the image read 0 (the reference reads whatever follows dev_depth_; parity rule = 0).
    const size_t idx = (size_t)x * (size_t)S.W + (size_t)y;
    const float d = idx < (size_t)S.W * (size_t)S.H ? __ldg(&depth[idx]) : 0.0f;
    if (d == 0.0f || d == __int_as_float(0xff800000)) R.alive = false;            // tsdf.cu:2116
    else if (d >= S.max_depth) R.alive = false;                                   // tsdf.cu:2119
    else if (fminf(S.max_depth, fsub(d, S.trunc)) >= fminf(S.max_depth, fadd(d, S.trunc))) R.alive = false;   // tsdf.cu:2122-2126
  }
  if (R.alive) {
    const Float3 r0 = pixel_to_world(F.c2w, S.fx, S.fy, S.cx, S.cy, (int)x, (int)y, S.min_depth);   // tsdf.cu:2129
    const Float3 r1 = pixel_to_world(F.c2w, S.fx, S.fy, S.cx, S.cy, (int)x, (int)y, S.max_depth);   // tsdf.cu:2130
    const float vx = fsub(r1.x, r0.x), vy = fsub(r1.y, r0.y), vz = fsub(r1.z, r0.z);
    const float inv = fdiv(1.0f, fsqrt(fadd(fadd(fmul(vx, vx), fmul(vy, vy)), fmul(vz, vz))));       // normalize, cutil_math.h:1207
    const float dir[3] = {fmul(vx, inv), fmul(vy, inv), fmul(vz, inv)};
    const float rm[3] = {r0.x, r0.y, r0.z}, rM[3] = {r1.x, r1.y, r1.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      R.cur[a] = __float2int_rd(fdiv(rm[a], S.block_size));                                           // tsdf.cu:2136
      const int end = __float2int_rd(fdiv(rM[a], S.block_size));                                      // tsdf.cu:2137
      const float st = sign_f(dir[a]);                                                                // tsdf.cu:2145
      R.istep[a] = (int)st;
      const int up = st > 0.0f ? 1 : 0;                                                               // clamp(step, 0, 1)
      const float boundary = fsub(fmul(i2f(R.cur[a] + up), S.block_size), S.half_vox);                // tsdf.cu:2146
      R.tmax[a] = fdiv(fsub(boundary, rm[a]), dir[a]);                                                // tsdf.cu:2147
      R.tdel[a] = fdiv(fmul(fmul(st, S.vox_size), (float)VPB), dir[a]);                               // tsdf.cu:2148
      R.bound[a] = __float2int_rz(fadd(i2f(end), st));                                                // tsdf.cu:2149
      if (dir[a] == 0.0f || fsub(boundary, rm[a]) == 0.0f) { R.tmax[a] = __int_as_float(0x7f800000); R.tdel[a] = __int_as_float(0x7f800000); }
    }
  }
}

// march `steps` DDA steps of one ray, one key per step into out[step * RAYS]
__device__ __forceinline__ void ray_march(RayState& R, int steps, u64* __restrict__ out) {
  for (int it = 0; it < steps; ++it) {                                                                // tsdf.cu:2158
    u64 k = KEY_EMPTY;
    if (R.alive) {
      if (key_in_range(R.cur[0], R.cur[1], R.cur[2])) k = pack_key(R.cur[0], R.cur[1], R.cur[2]);
      // advance (tsdf.cu:2217-2233); the block entered when a ray reaches its bound is never tested
      if (R.tmax[0] < R.tmax[1] && R.tmax[0] < R.tmax[2]) {
        R.cur[0] += R.istep[0]; if (R.cur[0] == R.bound[0]) R.alive = false; R.tmax[0] = fadd(R.tmax[0], R.tdel[0]);
      } else if (R.tmax[2] < R.tmax[1]) {
        R.cur[2] += R.istep[2]; if (R.cur[2] == R.bound[2]) R.alive = false; R.tmax[2] = fadd(R.tmax[2], R.tdel[2]);
      } else {
        R.cur[1] += R.istep[1]; if (R.cur[1] == R.bound[1]) R.alive = false; R.tmax[1] = fadd(R.tmax[1], R.tdel[1]);
      }
    }
    out[it * RAYS] = k;
  }
}

// Pipeline per CTA: warp 0 (8 lanes) marches the rays CHUNK_STEPS steps at a time into a double-buffered key list while
// warps 1-7 classify, insert and stamp the previous stage's keys; first-seen blocks are collected in shared memory and
// flushed to the visible list with ONE global atomicAdd per CTA (the per-warp atomics on the single counter used to
// serialise in L2: 24 % of the stall samples).
__global__ void __launch_bounds__(ALLOC_THREADS)
alloc_visible_kernel(const StaticParams S, const FrameParams F, const float* __restrict__ depth, const DeviceView D, int tiles_x) {
#ifdef VH_HOST_EMU                                   // CPU emulation of the kernel sources (tests/emu, test infrastructure)
  u64* dyn = reinterpret_cast<u64*>(emu::g_cta->dyn_smem);
#else
  extern __shared__ u64 dyn[];                       // [2][CHUNK_KEYS] keys, then int first[max_steps * RAYS]
#endif
  u64* skeys = dyn;
  int* s_first = reinterpret_cast<int*>(dyn + 2 * CHUNK_KEYS);
  __shared__ int s_cnt, s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int nchunks = (S.max_steps + CHUNK_STEPS - 1) / CHUNK_STEPS;
  if (tid == 0) s_cnt = 0;

  RayState R;
  R.alive = false;
  if (tid < RAYS) {
    ray_setup(S, F, depth, tile_x * RAYS_X + (tid & (RAYS_X - 1)), tile_y * RAYS_Y + (tid / RAYS_X), R);
    ray_march(R, min(CHUNK_STEPS, S.max_steps), skeys + tid);
  }
  __syncthreads();

  const float bpc = (float)S.bpc;
  for (int c = 0; c < nchunks; ++c) {
    if (wid == 0) {
      if (tid < RAYS && c + 1 < nchunks) ray_march(R, min(CHUNK_STEPS, S.max_steps - (c + 1) * CHUNK_STEPS), skeys + ((c + 1) & 1) * CHUNK_KEYS + tid);
    } else {
      const int i = tid - 32;
      const int nkeys = min(CHUNK_STEPS, S.max_steps - c * CHUNK_STEPS) * RAYS;
      u64 key = i < nkeys ? skeys[(c & 1) * CHUNK_KEYS + i] : KEY_EMPTY;
      if (key != KEY_EMPTY) {
        int bx, by, bz;
        unpack_key(key, bx, by, bz);
        bool ok = chunk_is_candidate(S, F, block_to_chunk(bx, bpc), block_to_chunk(by, bpc), block_to_chunk(bz, bpc));   // tsdf.cu:2164
        if (ok) ok = block_in_frustum(S, F, bx, by, bz);                                                                  // tsdf.cu:2165
        if (ok && S.shard_count > 1) ok = owner_of_block(bx, by, bz, S.shard_count, S.shard_group) == S.shard_rank;
        if (!ok) key = KEY_EMPTY;
      }
      if (__ballot_sync(0xffffffffu, key != KEY_EMPTY) != 0) {
        const unsigned same = __match_any_sync(0xffffffffu, key);
        const bool leader = key != KEY_EMPTY && lane == __ffs(same) - 1;
        int entry = -1;
        bool claimed = false;
        if (leader) entry = map_claim(D.map, key, claimed);
        map_assign_slots(D.map, 0xffffffffu, claimed, entry, key);
        bool first = false;
        if (leader && entry >= 0) first = atomicExch(&D.stamps[entry], F.frame) != F.frame;
        const unsigned fm = __ballot_sync(0xffffffffu, first);
        if (fm) {
          const int l0 = __ffs(fm) - 1;
          int base = 0;
          if (lane == l0) base = atomicAdd(&s_cnt, __popc(fm));
          base = __shfl_sync(0xffffffffu, base, l0);
          if (first) s_first[base + __popc(fm & ((1u << lane) - 1))] = entry;
        }
      }
    }
    __syncthreads();
  }
  // flush: one reservation in the global visible list per CTA
  const int cnt = s_cnt;
  if (cnt == 0) return;
  if (tid == 0) s_base = atomicAdd(&D.counters->visible_count, cnt);
  __syncthreads();
  const int base = s_base;
  for (int i = tid; i < cnt; i += ALLOC_THREADS)
    if (base + i < D.list_cap) D.visible[base + i] = s_first[i];
}

// ======================================================================================================================
// The 3-D DDA as a MERGE (used by ray_keys_kernel below). The reference's loop (tsdf.cu:2158-2233) keeps one crossing time per
// axis and advances the axis with the smallest one, adding that axis's increment by repeated float addition. The k-th crossing
// time of an axis therefore does not depend on the other axes: T_a[k] = tmax_a + tdel_a + ... (k additions) is a monotone
// sequence that one lane can generate alone (a chain of K dependent FADDs, bit-identical to the reference's accumulation), and
// the loop's choice
//     x if tx < ty && tx < tz;  else z if tz < ty;  else y
// is "take the smallest head, ties resolved y before z before x" — a three-way MERGE of the sequences under the total
// order (time, priority). Every element then finds its own place: element k of axis a is step
//     pos = k + #{elements of b before it} + #{elements of c before it}
// and those same counts are how far the ray has moved along b and c when it takes that step — so the block visited at step
// pos is (cur0_a + k*step_a, cur0_b + n_b*step_b, cur0_c + n_c*step_c). The ray ends at the first step that carries a
// coordinate onto its bound (an atomicMin over the three candidates). The sequential march (~40 dependent instructions per
// step) becomes parallel work over (ray, axis, k). Non-finite crossing times (a pose with NaNs) cannot index out of range:
// positions are checked, unwritten steps stay empty. Under CPU emulation the merge equals the step-by-step DDA key by key, exact
// ties between axes included (tests/test_emu_engine.py); on B200 it reproduces the oracle's visible sets on every config.
constexpr int ALLOC1_MAX_SMEM = 224 * 1024;      // of the 227 KB a CTA can have: ray_keys_kernel's smallest tile (2 rays, 40 B per step) handles step caps up to ~5,700

__device__ __forceinline__ int axis_priority(int a) { return a == 1 ? 0 : (a == 2 ? 1 : 2); }   // ties: y, then z, then x

// number of elements of the monotone sequence t[0..n) that precede the value v (ties_first: equal elements precede it too)
__device__ __forceinline__ int merge_rank(const float* __restrict__ t, int n, float v, bool ties_first) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const float x = t[mid];
    if (x < v || (ties_first && x == v)) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// The same rank, found from where it is expected: the sequence is t0 plus j increments up to rounding, so (v - t0) / del
// lands within an element of the boundary; a walk of at most three elements either way settles the exact place (the
// answer is defined by the comparisons with the stored elements, never by the estimate), and anything else — a direction
// component of zero (all elements +inf), non-finite values — falls back to the bisection above.
__device__ __forceinline__ int merge_rank_near(const float* __restrict__ t, int n, float v, bool ties_first, float inv_del) {
  const float p = __fmul_rn(__fsub_rn(v, t[0]), inv_del);
  int j = p >= 0.0f ? (p < (float)n ? (int)p + 1 : n) : 0;               // NaN -> 0
#pragma unroll 1
  for (int tries = 0; tries < 4; tries++) {
    if (j > 0) { const float x = t[j - 1]; if (!(x < v || (ties_first && x == v))) { j--; continue; } }
    if (j < n) { const float x = t[j]; if (x < v || (ties_first && x == v)) { j++; continue; } }
    return j;
  }
  return merge_rank(t, n, v, ties_first);
}

// ======================================================================================================================
// Allocation as TWO kernels with a key list in between — ray_keys_kernel -> (inbox) -> insert_keys_kernel (alloc_rev 2: sharded maps
// and long ray step caps; the one-kernel form above stays the default for the reference's 100-step cap on one GPU, where it is
// ~5 us faster: profiles/r02e). The one-kernel form interleaves a few thousand DDA steps with hash-table insertions inside every CTA: each round of
// 256 keys walks a chain of dependent global accesses (probe -> CAS -> pool pop -> stamp exchange -> list append) with at most
// 20 warps per SM to hide it (profiles/r02b: 24 us, issue slots 28 % busy, barrier + long-scoreboard stalls), and at the
// 1,100-step cap of the room-scale config that chain repeats 34 times per CTA (0.40 ms). Split in two:
//   ray_keys_kernel    the merge formulation of the DDA (above) for a tile of TRX x TRY rays, then every key of
//                      the tile is classified (ray end, chunk candidate test, frustum test) — pure arithmetic, no global access — and
//                      the survivors are appended to the INBOX of the GPU that owns the block: one reservation per owner per CTA.
//                      On a single GPU the owner is this GPU; on a sharded map the inbox of a peer is written over NVLink (stores
//                      into its mapped memory), and each GPU marches only every shard_count-th tile of rays: the ray pass is SPLIT
//                      across the GPUs, not replicated.
//   insert_keys_kernel one thread per inbox key, as many CTAs as the GPU holds: claim (64-bit CAS), pool pop and key_heap append
//                      aggregated per CTA (ONE atomicSub + ONE atomicAdd per 512 keys: at ~1 M new blocks per frame per-warp atomics
//                      on the two counters would serialise in L2), stamp exchange, first-seen entries appended to the visible list
//                      with one atomicAdd per CTA.
// Between the two a sharded map needs every peer's keys to have arrived: the frame barrier of vh_shard.cu. The inbox is
// double-buffered by frame parity (a peer may already be writing frame f+1 while this GPU inserts frame f).
template <int TRX, int TRY>
__device__ __forceinline__ void merge_fill_keys_t(const StaticParams& S, const FrameParams& F, const float* __restrict__ depth, int tile_x, int tile_y,
                                                  u64* __restrict__ skeys, float* __restrict__ sT, int* __restrict__ s_death) {
  constexpr int NR = TRX * TRY;
  __shared__ int s_cur[NR][3], s_step[NR][3], s_last[NR][3], s_alive[NR];
  __shared__ float s_del[NR][3], s_inv[NR][3];
  const int K = S.max_steps;
  const int tid = threadIdx.x, nthreads = (int)blockDim.x;

  // phase 0: ray set-up (one lane per ray), every step slot empty
  if (tid < NR) {
    RayState R;
    ray_setup(S, F, depth, tile_x * TRX + (tid % TRX), tile_y * TRY + (tid / TRX), R);
    s_alive[tid] = R.alive ? 1 : 0;
    s_death[tid] = K;                                                 // no step ends the ray (yet)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      s_cur[tid][a] = R.cur[a]; s_step[tid][a] = R.istep[a];
      // the a-step that carries cur_a onto bound_a is number (bound - cur) / step, counted from 1; none if the bound is not ahead
      const long long ahead = ((long long)R.bound[a] - (long long)R.cur[a]) * (long long)R.istep[a];
      s_last[tid][a] = (R.istep[a] != 0 && ahead >= 1 && ahead <= (long long)K) ? (int)ahead - 1 : -1;
      sT[(tid * 3 + a) * K] = R.tmax[a];
      s_del[tid][a] = R.tdel[a];
      s_inv[tid][a] = R.tdel[a] > 0.0f ? fdiv(1.0f, R.tdel[a]) : 0.0f;         // estimate only (merge_rank_near); +inf -> 0
    }
  }
  for (int i = tid; i < K * NR; i += nthreads) skeys[i] = KEY_EMPTY;
  __syncthreads();

  // phase 1: the crossing times of every axis by repeated addition (tsdf.cu:2221,2226,2231), one lane per (ray, axis)
  if (tid < NR * 3) {
    float* t = sT + (size_t)tid * K;
    float v = t[0];
    const float del = s_del[tid / 3][tid % 3];
    for (int k = 1; k < K; k++) { v = fadd(v, del); t[k] = v; }
  }
  __syncthreads();

  // phase 2: every element finds its step and the block the ray is in when it takes it: items (ray, axis, k) flattened over the CTA
  for (int i = tid; i < NR * 3 * K; i += nthreads) {
    const int ra = i / K, k = i - ra * K;
    const int ray = ra / 3, a = ra - ray * 3;
    if (!s_alive[ray]) continue;
    const int b = a == 0 ? 1 : 0, c = a == 2 ? 1 : 2;                 // the other two axes
    const int pa = axis_priority(a);
    const bool tb = axis_priority(b) < pa, tc = axis_priority(c) < pa;
    const float* Tb = sT + (size_t)(ray * 3 + b) * K;
    const float* Tc = sT + (size_t)(ray * 3 + c) * K;
    const float v = sT[(size_t)ra * K + k];
    const int nb = merge_rank_near(Tb, K, v, tb, s_inv[ray][b]);
    if (k + nb >= K) continue;
    const int nc = merge_rank_near(Tc, K, v, tc, s_inv[ray][c]);
    const int pos = k + nb + nc;
    if (pos < 0 || pos >= K) continue;
    int cur[3];
    cur[a] = s_cur[ray][a] + k * s_step[ray][a]; cur[b] = s_cur[ray][b] + nb * s_step[ray][b]; cur[c] = s_cur[ray][c] + nc * s_step[ray][c];
    if (key_in_range(cur[0], cur[1], cur[2])) skeys[pos * NR + ray] = pack_key(cur[0], cur[1], cur[2]);
    if (k == s_last[ray][a]) atomicMin(&s_death[ray], pos);           // this step moves the ray onto its bound (tsdf.cu:2219,2224,2229)
  }
  __syncthreads();
}

constexpr int KEYS_THREADS = 256;
template <int TRX, int TRY>
__global__ void __launch_bounds__(KEYS_THREADS)
ray_keys_kernel(const __grid_constant__ StaticParams S, const __grid_constant__ FrameParams F, const float* __restrict__ depth, const __grid_constant__ DeviceView D,
                int tiles_x, int n_tiles) {
  constexpr int NR = TRX * TRY;
#ifdef VH_HOST_EMU
  u64* dyn = reinterpret_cast<u64*>(emu::g_cta->dyn_smem);
#else
  extern __shared__ u64 dyn[];
#endif
  const int K = S.max_steps;
  u64* skeys = dyn;                                                   // [K][NR]
  float* sT = reinterpret_cast<float*>(dyn + (size_t)K * NR);         // [NR][3][K] crossing times (dead after phase 2)
  __shared__ int s_death[NR];
  __shared__ int s_cnt[MAX_SHARDS], s_base[MAX_SHARDS], s_fill[MAX_SHARDS];
  const int tid = threadIdx.x;
  // a connected sharded map splits the tiles of rays across its GPUs (keys travel to their owner's inbox); a shard without
  // peers marches every tile and keeps its own keys
  const bool routed = S.shard_count > 1 && D.peers != nullptr;
  const int tile = routed ? (int)blockIdx.x * (int)S.shard_count + (int)S.shard_rank : (int)blockIdx.x;
  if (tile >= n_tiles) return;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  if (tid < MAX_SHARDS) { s_cnt[tid] = 0; s_fill[tid] = 0; }
  merge_fill_keys_t<TRX, TRY>(S, F, depth, tile_x, tile_y, skeys, sT, s_death);

  // classify every key of the tile; the survivors stay in skeys (the rest become KEY_EMPTY) and are counted per owner
  const float bpc = (float)S.bpc;
  const int nkeys = K * NR;
  const int parity = (int)(F.frame & 1u);
  for (int i = tid; i < nkeys; i += KEYS_THREADS) {
    u64 key = skeys[i];
    if (key == KEY_EMPTY) continue;
    bool ok = i / NR <= s_death[i % NR];
    int owner = 0;
    if (ok) {
      int bx, by, bz;
      unpack_key(key, bx, by, bz);
      ok = chunk_is_candidate(S, F, block_to_chunk(bx, bpc), block_to_chunk(by, bpc), block_to_chunk(bz, bpc));   // tsdf.cu:2164
      if (ok) ok = block_in_frustum(S, F, bx, by, bz);                                                          // tsdf.cu:2165
      if (ok && S.shard_count > 1) {
        owner = (int)owner_of_block(bx, by, bz, S.shard_count, S.shard_group);
        if (!routed) { ok = owner == (int)S.shard_rank; owner = 0; }
      }
    }
    if (ok) atomicAdd(&s_cnt[owner], 1); else skeys[i] = KEY_EMPTY;
  }
  __syncthreads();
  // one reservation per owner per CTA, in the owner's inbox (a peer's: a remote atomic over NVLink)
  if (tid < MAX_SHARDS && s_cnt[tid] > 0) {
    int* cnt = routed ? D.peers->v[tid].inbox_count : D.inbox_count;
    s_base[tid] = routed ? atomicAdd_system(&cnt[parity], s_cnt[tid]) : atomicAdd(&cnt[parity], s_cnt[tid]);   // a peer's counter: system scope
  }
  __syncthreads();
  for (int i = tid; i < nkeys; i += KEYS_THREADS) {
    const u64 key = skeys[i];
    if (key == KEY_EMPTY) continue;
    int owner = 0;
    if (routed) { int bx, by, bz; unpack_key(key, bx, by, bz); owner = (int)owner_of_block(bx, by, bz, S.shard_count, S.shard_group); }
    const int pos = s_base[owner] + atomicAdd(&s_fill[owner], 1);
    u64* box = routed ? D.peers->v[owner].inbox : D.inbox;
    if (pos < D.inbox_cap) box[(size_t)parity * D.inbox_cap + pos] = key;
    else atomicOr(D.map.error_flag, MAP_TABLE_FULL);
  }
}

inline size_t ray_keys_smem_bytes(int max_steps, int rays) {
  return (size_t)max_steps * rays * sizeof(u64) + (size_t)rays * 3 * max_steps * sizeof(float);     // keys + crossing times (20 B per ray-step)
}

constexpr int INSERT_THREADS = 512;
__global__ void __launch_bounds__(INSERT_THREADS)
insert_keys_kernel(const __grid_constant__ DeviceView D, const uint32_t frame) {
  __shared__ int s_wclaim[INSERT_THREADS / 32], s_wfirst[INSERT_THREADS / 32];
  __shared__ int s_top, s_heap, s_vis;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int parity = (int)(frame & 1u);
  const int n = min(D.inbox_count[parity], D.inbox_cap);
  const u64* box = D.inbox + (size_t)parity * D.inbox_cap;
  for (int base = (int)blockIdx.x * INSERT_THREADS; base < n; base += (int)gridDim.x * INSERT_THREADS) {
    const int i = base + tid;
    const u64 key = i < n ? box[i] : KEY_EMPTY;
    int entry = -1;
    bool claimed = false;
    if (key != KEY_EMPTY) entry = map_claim(D.map, key, claimed);
    bool first = false;
    if (entry >= 0) first = atomicExch(&D.stamps[entry], frame) != frame;     // exactly one thread per block and frame
    const unsigned cm = __ballot_sync(0xffffffffu, claimed), fm = __ballot_sync(0xffffffffu, first);
    if (lane == 0) { s_wclaim[wid] = __popc(cm); s_wfirst[wid] = __popc(fm); }
    __syncthreads();
    if (wid == 0) {     // exclusive scan of the per-warp counts; ONE pop, ONE key_heap reservation, ONE visible-list reservation per CTA
      const int c = lane < INSERT_THREADS / 32 ? s_wclaim[lane] : 0, f = lane < INSERT_THREADS / 32 ? s_wfirst[lane] : 0;
      int ci = c, fi = f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ci, o), v = __shfl_up_sync(0xffffffffu, fi, o); if (lane >= o) { ci += u; fi += v; } }
      if (lane < INSERT_THREADS / 32) { s_wclaim[lane] = ci - c; s_wfirst[lane] = fi - f; }
      const int nc = __shfl_sync(0xffffffffu, ci, 31), nf = __shfl_sync(0xffffffffu, fi, 31);
      if (lane == 0) {          // pool pop, then the key_heap reservation for what was granted
        int top = 0, heap = 0;
        if (nc > 0) {
          top = atomicSub(D.map.free_top, nc);
          const int granted = top >= nc ? nc : (top > 0 ? top : 0);
          if (granted < nc) atomicAdd(D.map.free_top, nc - granted);          // pool exhausted: hand back the share that was not there
          if (granted > 0) heap = atomicAdd(D.map.heap_counter, granted);
        }
        s_top = top; s_heap = heap;
      } else if (lane == 1) {   // the visible-list reservation travels at the same time
        s_vis = nf > 0 ? atomicAdd(&D.counters->visible_count, nf) : 0;
      }
    }
    __syncthreads();
    if (claimed) {
      const int rank = s_wclaim[wid] + __popc(cm & ((1u << lane) - 1));
      const int idx = s_top - 1 - rank;
      if (idx >= 0) {
        D.map.slots[entry] = D.map.free_list[idx];
        if (s_heap + rank < D.map.num_blocks) D.map.key_heap[s_heap + rank] = key;
      } else {
        D.map.slots[entry] = SLOT_POOL_FULL;
        atomicOr(D.map.error_flag, MAP_POOL_FULL);
      }
    }
    if (first) {
      const int pos = s_vis + s_wfirst[wid] + __popc(fm & ((1u << lane) - 1));
      if (pos < D.list_cap) D.visible[pos] = entry;
    }
    __syncthreads();                 // the shared counters are rewritten in the next round
  }
  // The last CTA to finish empties the inbox of this parity for frame + 2 (every CTA read the count before it got here; a
  // peer writes this parity again only after the next frame's barrier, which this GPU reaches after this kernel).
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&D.inbox_done[parity], 1) == (int)gridDim.x - 1) { D.inbox_count[parity] = 0; D.inbox_done[parity] = 0; __threadfence(); }
  }
}

#ifndef VH_HOST_EMU
// first half of the allocation: the frame's block keys into their owners' inboxes (revision 2)
template <int TRX, int TRY>
static void launch_ray_keys_t(const StaticParams& S, const FrameParams& F, const float* d_depth, const DeviceView& D, cudaStream_t st) {
  const int tiles_x = (S.nrx + TRX - 1) / TRX, tiles_y = (S.nry + TRY - 1) / TRY, n_tiles = tiles_x * tiles_y;
  const bool routed = S.shard_count > 1 && D.peers != nullptr;
  const int grid = routed ? (n_tiles + (int)S.shard_count - 1) / (int)S.shard_count : n_tiles;
  const size_t smem = ray_keys_smem_bytes(S.max_steps, TRX * TRY);
  if (smem > 48 * 1024) cudaFuncSetAttribute(ray_keys_kernel<TRX, TRY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ray_keys_kernel<TRX, TRY><<<grid, KEYS_THREADS, smem, st>>>(S, F, d_depth, D, tiles_x, n_tiles);
}
bool alloc_uses_inbox(const StaticParams& S) { return S.alloc_rev == 2 && ray_keys_smem_bytes(S.max_steps, 2) <= (size_t)ALLOC1_MAX_SMEM; }
void launch_ray_keys(const StaticParams& S, const FrameParams& F, const float* d_depth, const DeviceView& D, int num_sms, cudaStream_t st) {
  // the largest tile of rays that still gives every SM a few CTAs (a sharded map marches 1/shard_count of the tiles per GPU)
  const int rays = S.nrx * S.nry, share = (S.shard_count > 1 && D.peers) ? (int)S.shard_count : 1;
  if (rays / 8 / share >= 4 * num_sms && ray_keys_smem_bytes(S.max_steps, 8) <= 64 * 1024) launch_ray_keys_t<4, 2>(S, F, d_depth, D, st);
  else if (rays / 4 / share >= 4 * num_sms && ray_keys_smem_bytes(S.max_steps, 4) <= 96 * 1024) launch_ray_keys_t<2, 2>(S, F, d_depth, D, st);
  else launch_ray_keys_t<2, 1>(S, F, d_depth, D, st);
}
// second half: this GPU's inbox into its table, pool and visible list
void launch_insert_keys(const StaticParams& S, const FrameParams& F, const DeviceView& D, int num_sms, cudaStream_t st) {
  insert_keys_kernel<<<num_sms * 3, INSERT_THREADS, 0, st>>>(D, F.frame);
}

void launch_alloc_visible(const StaticParams& S, const FrameParams& F, const float* d_depth, const DeviceView& D, cudaStream_t st) {
  if (alloc_uses_inbox(S)) {      // a connected sharded map needs its frame barrier between the two: vh_shard.cu launches them itself
    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    launch_ray_keys(S, F, d_depth, D, num_sms, st);
    launch_insert_keys(S, F, D, num_sms, st);
    return;
  }
  const int tiles_x = (S.nrx + RAYS_X - 1) / RAYS_X, tiles_y = (S.nry + RAYS_Y - 1) / RAYS_Y;
  const size_t smem = 2 * CHUNK_KEYS * sizeof(u64) + (size_t)S.max_steps * RAYS * sizeof(int);
  if (smem > 48 * 1024)   // per-device attribute; only very long ray step caps get here
    cudaFuncSetAttribute(alloc_visible_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  alloc_visible_kernel<<<tiles_x * tiles_y, ALLOC_THREADS, smem, st>>>(S, F, d_depth, D, tiles_x);
}
#endif  // !VH_HOST_EMU

// ---- caller-supplied visible list (stage tests; vh_set_visible) ------------------------------------
__global__ void set_visible_kernel(const DeviceView D, const u64* __restrict__ keys, int n, uint32_t frame) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  u64 key = i < n ? keys[i] : KEY_EMPTY;
  const unsigned same = __match_any_sync(0xffffffffu, key);
  const bool leader = key != KEY_EMPTY && lane == __ffs(same) - 1;
  int entry = -1;
  bool claimed = false;
  if (leader) entry = map_claim(D.map, key, claimed);
  map_assign_slots(D.map, 0xffffffffu, claimed, entry, key);
  bool first = false;
  if (leader && entry >= 0) first = atomicExch(&D.stamps[entry], frame) != frame;
  const unsigned fm = __ballot_sync(0xffffffffu, first);
  if (fm) {
    const int l0 = __ffs(fm) - 1;
    int base = 0;
    if (lane == l0) base = atomicAdd(&D.counters->visible_count, __popc(fm));
    base = __shfl_sync(0xffffffffu, base, l0);
    if (first) {
      const int pos = base + __popc(fm & ((1u << lane) - 1));
      if (pos < D.list_cap) D.visible[pos] = entry;
    }
  }
}

#ifndef VH_HOST_EMU
void launch_set_visible(const DeviceView& D, const u64* d_keys, int n, uint32_t frame, cudaStream_t st) {
  if (n <= 0) return;
  set_visible_kernel<<<(n + 255) / 256, 256, 0, st>>>(D, d_keys, n, frame);
}
#endif  // !VH_HOST_EMU

// every allocated block, in key_heap order (full-map marching cubes, exports)
__global__ void list_all_blocks_kernel(const DeviceView D, int* __restrict__ list, int* __restrict__ list_count) {
  const int n = min(*D.map.heap_counter, D.map.num_blocks);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *list_count = n;
  if (i < n) list[i] = map_find(D.map, D.map.key_heap[i]);
}

#ifndef VH_HOST_EMU
void launch_list_all_blocks(const DeviceView& D, int* list, int* list_count, cudaStream_t st) {
  const int n = D.map.num_blocks;
  list_all_blocks_kernel<<<(n + 255) / 256, 256, 0, st>>>(D, list, list_count);
}
#endif  // !VH_HOST_EMU

}  // namespace vh
