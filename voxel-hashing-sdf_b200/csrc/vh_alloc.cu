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

#ifndef VH_HOST_EMU
void launch_alloc_visible(const StaticParams& S, const FrameParams& F, const float* d_depth, const DeviceView& D, cudaStream_t st) {
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
