// vh_stream.cu — the out-of-core tier: whole voxel blocks move between the device map and caller-owned host memory.
//
// The reference keeps its map on the HOST (h_chunks) and streams it through the GPU every frame: streamInCPU2GPU uploads
// every chunk near the frustum centre (/root/reference/src/tsdf.cu:277-457), streamOutGPU2CPU copies the working set
// back (tsdf.cu:469-596). This engine keeps the map resident in HBM, so nothing moves on the hot
// path; this file is the optional tier for scenes beyond one GPU's memory, as three explicit
// calls made BETWEEN frames:
//   vh_far_blocks     which allocated blocks lie in chunks outside the reference's residency rule for a pose
//                     (chunk cube + chunk sphere around the frustum centre, tsdf.cu:166-187,300-312)
//   vh_evict_blocks   download those blocks and release their table entries and pool slots
//   vh_upload_blocks  insert blocks (back) with their voxels
// The host store and the policy stay with the caller (include/tsdf.cuh keeps the reference's chunk store on top of them).
// Released entries become tombstones (probe sequences run through them, insertion does not reuse them); when they exceed
// a quarter of the table the table is rebuilt in place from the live entries.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "vh_engine_host.h"
#include "vh_math.cuh"
#include "vh_params_host.h"

namespace vh {

// one warp per key: release the entry and the pool slot, leave the slot zeroed like a fresh one (reset_map)
__global__ void evict_kernel(const DeviceView D, const u64* __restrict__ keys, int n, int* __restrict__ released) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  int slot = -1;
  if (lane == 0) {
    const u64 key = keys[w];
    const int e = map_find(D.map, key);
    // the CAS makes a key that is listed twice release its slot once
    if (e >= 0 && atomicCAS(&D.map.keys[e], key, KEY_TOMB) == key) {
      slot = D.map.slots[e];
      D.map.slots[e] = -1;
      D.stamps[e] = 0u;
    }
  }
  slot = __shfl_sync(0xffffffffu, slot, 0);
  if (slot < 0) return;
  for (int v = lane; v < BLOCK_VOX; v += 32) {
    const size_t i = (size_t)slot * BLOCK_VOX + v;
    D.sdf[i] = 0.0f; D.wgt[i] = 0.0f;
    if (D.rgb) D.rgb[i] = make_uchar4(0, 0, 0, 0);
  }
  if (lane == 0) {
    D.neg_count[slot] = 0; D.tri_count[slot] = 0; D.tri_offset[slot] = 0ull;
    map_release_slot(D.map, slot);                                 // push: the stack only grows here, nothing pops concurrently
    atomicAdd(released, 1);
  }
}

// insert-if-absent for a list of keys (no stamping, no visible list): out_slots[i] = the block's pool slot, -1 on exhaustion
__global__ void upload_insert_kernel(const DeviceView D, const u64* __restrict__ keys, int n, int* __restrict__ out_slots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const u64 key = i < n ? keys[i] : KEY_EMPTY;
  const unsigned same = __match_any_sync(0xffffffffu, key);
  const int lead = __ffs(same) - 1;
  const bool leader = key != KEY_EMPTY && lane == lead;
  int entry = -1;
  bool claimed = false;
  if (leader) entry = map_claim(D.map, key, claimed);
  map_assign_slots(D.map, 0xffffffffu, claimed, entry, key);
  int slot = -1;
  if (leader && entry >= 0) slot = D.map.slots[entry];
  slot = __shfl_sync(0xffffffffu, slot, lead);
  if (i < n) out_slots[i] = slot;
}

// one warp per block: planes from the staging buffers into the slot, negative-voxel counter recomputed, no mesh yet
__global__ void upload_scatter_kernel(const DeviceView D, const int* __restrict__ slots, int n, const float* __restrict__ sdf, const float* __restrict__ wgt,
                                      const uint8_t* __restrict__ rgb) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int slot = slots[w];
  if (slot < 0) return;
  int neg = 0;
  for (int v = lane; v < BLOCK_VOX; v += 32) {
    const size_t src = (size_t)w * BLOCK_VOX + v, dst = (size_t)slot * BLOCK_VOX + v;
    const float s = sdf[src];
    D.sdf[dst] = s; D.wgt[dst] = wgt[src];
    neg += s < 0.0f ? 1 : 0;
    if (D.rgb) D.rgb[dst] = rgb ? make_uchar4(rgb[3 * src], rgb[3 * src + 1], rgb[3 * src + 2], 0) : make_uchar4(0, 0, 0, 0);
  }
  for (int o = 16; o > 0; o >>= 1) neg += __shfl_xor_sync(0xffffffffu, neg, o);
  if (lane == 0) { D.neg_count[slot] = neg; D.tri_count[slot] = 0; D.tri_offset[slot] = 0ull; }
}

// allocated blocks whose chunk fails the reference's residency rule for this frame's frustum centre
__device__ __forceinline__ bool chunk_resident(const StaticParams& S, const FrameParams& F, int x, int y, int z) {
  if (x < F.cstart[0] || x > F.cend[0] || y < F.cstart[1] || y > F.cend[1] || z < F.cstart[2] || z > F.cend[2]) return false;
  const double cs = (double)S.chunk_size;                                        // tsdf.cu:168: double arithmetic narrowed to float
  const float ccx = __double2float_rn(dmul(dadd((double)i2f(x), 0.5), cs));
  const float ccy = __double2float_rn(dmul(dadd((double)i2f(y), 0.5), cs));
  const float ccz = __double2float_rn(dmul(dadd((double)i2f(z), 0.5), cs));
  const float vx = fsub(F.fc[0], ccx), vy = fsub(F.fc[1], ccy), vz = fsub(F.fc[2], ccz);
  const float l = fsqrt(fadd(fadd(fmul(vx, vx), fmul(vz, vz)), fmul(vy, vy)));     // x, z, y order (tsdf.cu:174)
  return l <= fabsf(F.chunk_test_radius);
}
__global__ void far_blocks_kernel(const StaticParams S, const FrameParams F, const DeviceView D, u64* __restrict__ out, int cap, int* __restrict__ count) {
  const int n = min(*D.map.heap_counter, D.map.num_blocks);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 key = D.map.key_heap[i];
  int bx, by, bz;
  unpack_key(key, bx, by, bz);
  const float bpc = (float)S.bpc;
  const int cx = __float2int_rd(fdiv(i2f(bx), bpc)), cy = __float2int_rd(fdiv(i2f(by), bpc)), cz = __float2int_rd(fdiv(i2f(bz), bpc));   // tsdf.cu:256-260
  if (chunk_resident(S, F, cx, cy, cz)) return;
  const int pos = atomicAdd(count, 1);
  if (pos < cap) out[pos] = key;
}

// table rebuild: live entries out (compacted), table cleared, live entries back in
__global__ void rebuild_collect_kernel(const DeviceView D, uint32_t capacity, u64* __restrict__ k, int* __restrict__ s, uint32_t* __restrict__ st, int room,
                                       int* __restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= capacity) return;
  const u64 key = D.map.keys[i];
  if (key == KEY_EMPTY || key == KEY_TOMB) return;
  const int pos = atomicAdd(count, 1);
  if (pos < room) { k[pos] = key; s[pos] = D.map.slots[i]; st[pos] = D.stamps[i]; }
}
__global__ void rebuild_insert_kernel(const DeviceView D, const u64* __restrict__ k, const int* __restrict__ s, const uint32_t* __restrict__ st, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool claimed = false;
  const int e = map_claim(D.map, k[i], claimed);          // every key is distinct: each call claims its own entry
  if (e >= 0) { D.map.slots[e] = s[i]; D.stamps[e] = st[i]; }
}

}  // namespace vh

#ifndef VH_HOST_EMU
namespace {

int stage_keys(vh_engine* e, const int32_t* keys_xyz, int n, std::vector<u64>& packed) {
  packed.resize((size_t)n);
  for (int i = 0; i < n; i++) {
    const int x = keys_xyz[3 * i], y = keys_xyz[3 * i + 1], z = keys_xyz[3 * i + 2];
    if (!key_in_range(x, y, z)) return fail(VH_ERR_INVALID, "block coordinate (%d,%d,%d) outside [-2^20, 2^20)", x, y, z);
    packed[(size_t)i] = pack_key(x, y, z);
  }
  if ((size_t)n > e->keys_tmp_cap) {
    cudaFree(e->d_keys_tmp); e->d_keys_tmp = nullptr;
    CK(cudaMalloc((void**)&e->d_keys_tmp, (size_t)n * sizeof(u64)));
    e->keys_tmp_cap = (size_t)n;
  }
  CK(cudaMemcpyAsync(e->d_keys_tmp, packed.data(), (size_t)n * sizeof(u64), cudaMemcpyHostToDevice, e->stream));
  return VH_OK;
}

// the last frame's visible list holds entry indices: it is void once entries were released or the table was rebuilt
int invalidate_visible_and_refresh(vh_engine* e) {
  CK(cudaMemsetAsync(&e->D.counters->visible_count, 0, sizeof(int), e->stream));
  int rc = enqueue_readback(e);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return VH_OK;
}

int rebuild_table(vh_engine* e) {
  DeviceView& D = e->D;
  u64* k = nullptr; int* s = nullptr; uint32_t* st = nullptr; int* cnt = nullptr;
  // live entries: one per allocated block (plus entries left without a slot by an exhausted pool, an error state)
  const size_t nb = std::min<size_t>((size_t)e->capacity, 2 * (size_t)e->P.pool_blocks);
  CK(cudaMalloc((void**)&k, nb * sizeof(u64)));
  CK(cudaMalloc((void**)&s, nb * sizeof(int)));
  CK(cudaMalloc((void**)&st, nb * sizeof(uint32_t)));
  CK(cudaMalloc((void**)&cnt, sizeof(int)));
  cudaError_t ce = cudaMemsetAsync(cnt, 0, sizeof(int), e->stream);
  rebuild_collect_kernel<<<(e->capacity + 255) / 256, 256, 0, e->stream>>>(D, e->capacity, k, s, st, (int)nb, cnt);
  int n = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&n, cnt, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess && n > (int)nb) { cudaFree(k); cudaFree(s); cudaFree(st); cudaFree(cnt); return fail(VH_ERR_TABLE_FULL, "hash table rebuild: %d live entries for a pool of %d blocks", n, e->P.pool_blocks); }
  if (ce == cudaSuccess) ce = cudaMemsetAsync(D.map.keys, 0xFF, (size_t)e->capacity * sizeof(u64), e->stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(D.map.slots, 0xFF, (size_t)e->capacity * sizeof(int), e->stream);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(D.stamps, 0, (size_t)e->capacity * sizeof(uint32_t), e->stream);
  if (ce == cudaSuccess && n > 0) rebuild_insert_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(D, k, s, st, n);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(k); cudaFree(s); cudaFree(st); cudaFree(cnt);
  if (ce != cudaSuccess) return fail(VH_ERR_CUDA, "CUDA Error: %s while rebuilding the hash table", cudaGetErrorString(ce));
  e->tombstones = 0;
  return VH_OK;
}

}  // namespace

extern "C" {

int vh_far_blocks(vh_engine* e, const float* c2w, int32_t* out_xyz, int cap, int* n_out) {
  if (!e || !c2w || !n_out) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  FrameParams F;
  derive_frame_params(e->P, e->S, c2w, F);
  F.frame = 0;
  const int nb = e->P.pool_blocks;
  u64* d_out = nullptr; int* d_cnt = nullptr;
  const int room = out_xyz ? std::max(cap, 0) : 0;
  CK(cudaMalloc((void**)&d_out, (size_t)std::max(room, 1) * sizeof(u64)));
  CK(cudaMalloc((void**)&d_cnt, sizeof(int)));
  cudaError_t ce = cudaMemsetAsync(d_cnt, 0, sizeof(int), e->stream);
  far_blocks_kernel<<<(nb + 255) / 256, 256, 0, e->stream>>>(e->S, F, e->D, d_out, room, d_cnt);
  int n = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&n, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  std::vector<u64> keys((size_t)std::min(n, room));
  if (ce == cudaSuccess && !keys.empty()) ce = cudaMemcpy(keys.data(), d_out, keys.size() * sizeof(u64), cudaMemcpyDeviceToHost);
  cudaFree(d_out); cudaFree(d_cnt);
  if (ce != cudaSuccess) return fail(VH_ERR_CUDA, "CUDA Error: %s in vh_far_blocks", cudaGetErrorString(ce));
  std::sort(keys.begin(), keys.end());                    // the kernel appends in scheduling order: give the caller a stable one
  for (size_t i = 0; i < keys.size(); i++) unpack_key(keys[i], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
  *n_out = n;
  return VH_OK;
}

// The residency rule on the host, for blocks that are NOT on the device (the caller's store): plain float expressions in the
// order of chunk_resident above (this file is compiled with -ffp-contract=off), no engine or GPU needed.
int vh_blocks_resident(const vh_params* p, const float* c2w, const int32_t* keys_xyz, int n, uint8_t* out) {
  if (!p || !c2w || (n > 0 && (!keys_xyz || !out))) return fail(VH_ERR_INVALID, "null argument");
  if (p->vox_size <= 0 || p->blocks_per_chunk <= 0 || p->dda_stride <= 0) return fail(VH_ERR_INVALID, "invalid parameter value");
  StaticParams S; memset(&S, 0, sizeof(S));
  derive_static_params(*p, S);
  FrameParams F;
  derive_frame_params(*p, S, c2w, F);
  const float bpc = (float)S.bpc;
  const double cs = (double)S.chunk_size;
  for (int i = 0; i < n; i++) {
    int c[3];
    for (int a = 0; a < 3; a++) c[a] = (int)floorf((float)keys_xyz[3 * i + a] / bpc);                       // tsdf.cu:256-260
    bool in = true;
    for (int a = 0; a < 3; a++) in = in && c[a] >= F.cstart[a] && c[a] <= F.cend[a];
    if (in) {
      const float ccx = (float)(((double)(float)c[0] + 0.5) * cs), ccy = (float)(((double)(float)c[1] + 0.5) * cs), ccz = (float)(((double)(float)c[2] + 0.5) * cs);
      const float vx = F.fc[0] - ccx, vy = F.fc[1] - ccy, vz = F.fc[2] - ccz;
      const float l = sqrtf(vx * vx + vz * vz + vy * vy);                                                     // x, z, y order (tsdf.cu:174)
      in = l <= fabsf(F.chunk_test_radius);
    }
    out[i] = in ? 1 : 0;
  }
  return VH_OK;
}

int vh_evict_blocks(vh_engine* e, const int32_t* keys_xyz, int n, float* sdf, float* weight, uint8_t* rgb, uint8_t* found) {
  if (!e || (n > 0 && !keys_xyz)) return fail(VH_ERR_INVALID, "null argument");
  if (n <= 0) return VH_OK;
  if (e->shard) return fail(VH_ERR_INVALID, "the out-of-core tier is not available on a sharded map");
  // 1. the voxels, through the ordinary download path (takes the lock itself)
  int rc = vh_download_blocks(e, keys_xyz, n, sdf, weight, rgb, found);
  if (rc != VH_OK) return rc;
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  // 2. release entries and slots
  std::vector<u64> packed;
  rc = stage_keys(e, keys_xyz, n, packed);
  if (rc != VH_OK) return rc;
  int* d_rel = nullptr;
  CK(cudaMalloc((void**)&d_rel, sizeof(int)));
  cudaError_t ce = cudaMemsetAsync(d_rel, 0, sizeof(int), e->stream);
  evict_kernel<<<(n + 7) / 8, 256, 0, e->stream>>>(e->D, e->d_keys_tmp, n, d_rel);
  int released = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&released, d_rel, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(d_rel);
  if (ce != cudaSuccess) return fail(VH_ERR_CUDA, "CUDA Error: %s in vh_evict_blocks", cudaGetErrorString(ce));
  // 3. key_heap keeps the allocated keys in insertion order: drop the released ones (on the host: 8 bytes per block, next to the
  //    6 KB per block that were just downloaded)
  if (released > 0) {
    int heap_n = 0;
    CK(cudaMemcpy(&heap_n, e->D.map.heap_counter, sizeof(int), cudaMemcpyDeviceToHost));
    heap_n = std::min(heap_n, e->P.pool_blocks);
    std::vector<u64> heap((size_t)heap_n);
    if (heap_n) CK(cudaMemcpy(heap.data(), e->D.map.key_heap, (size_t)heap_n * sizeof(u64), cudaMemcpyDeviceToHost));
    const std::unordered_set<u64> gone(packed.begin(), packed.end());
    size_t m = 0;
    for (size_t i = 0; i < heap.size(); i++) if (!gone.count(heap[i])) heap[m++] = heap[i];
    const int new_n = (int)m;
    if (new_n != heap_n - released) return fail(VH_ERR_CUDA, "vh_evict_blocks: key list and table disagree (%d listed, %d released, %d left)", heap_n, released, new_n);
    if (new_n) CK(cudaMemcpy(e->D.map.key_heap, heap.data(), m * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->D.map.heap_counter, &new_n, sizeof(int), cudaMemcpyHostToDevice));
    e->tombstones += (uint32_t)released;
    if (e->tombstones > e->capacity / 4) { rc = rebuild_table(e); if (rc != VH_OK) return rc; }
  }
  return invalidate_visible_and_refresh(e);
}

int vh_upload_blocks(vh_engine* e, const int32_t* keys_xyz, int n, const float* sdf, const float* weight, const uint8_t* rgb) {
  if (!e || (n > 0 && (!keys_xyz || !sdf || !weight))) return fail(VH_ERR_INVALID, "null argument");
  if (n <= 0) return VH_OK;
  if (e->shard) return fail(VH_ERR_INVALID, "the out-of-core tier is not available on a sharded map");
  // The integrate kernel picks its exact short colour averages from an upper bound of every stored weight (launches since the
  // last reset). Uploaded blocks bring their own weights: they must be finite and non-negative, their maximum joins the
  // bound, and a weight that is not an integer (never produced by this engine or the reference) switches to the general
  // colour sequence for good, because the short forms are proven for integer weights only.
  float max_w = 0.0f; bool integral = true;
  for (size_t i = 0, m = (size_t)n * BLOCK_VOX; i < m; i++) {
    const float w = weight[i];
    if (!(w >= 0.0f) || !std::isfinite(w)) return fail(VH_ERR_INVALID, "vh_upload_blocks: weight[%zu] = %g is negative or not finite", i, (double)w);
    if (w > max_w) max_w = w;
    if (w != std::floor(w)) integral = false;
  }
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  if (!integral || max_w >= 16777216.0f) e->weight_bound_bias = 1u << 24;
  else e->integrate_launches = std::max<uint32_t>(e->integrate_launches, (uint32_t)max_w);
  const int CH = 16384;   // blocks per staging round (like vh_download_blocks)
  const int m0 = std::min(n, CH);
  float *d_s = nullptr, *d_w = nullptr; uint8_t* d_c = nullptr; int* d_slots = nullptr;
  CK(cudaMalloc((void**)&d_s, (size_t)m0 * BLOCK_VOX * sizeof(float)));
  CK(cudaMalloc((void**)&d_w, (size_t)m0 * BLOCK_VOX * sizeof(float)));
  if (rgb) CK(cudaMalloc((void**)&d_c, (size_t)m0 * BLOCK_VOX * 3));
  CK(cudaMalloc((void**)&d_slots, (size_t)m0 * sizeof(int)));
  int rc = VH_OK;
  std::vector<u64> packed;
  for (int o = 0; o < n && rc == VH_OK; o += CH) {
    const int m = std::min(CH, n - o);
    rc = stage_keys(e, keys_xyz + 3 * (size_t)o, m, packed);
    if (rc != VH_OK) break;
    cudaError_t ce = cudaMemcpyAsync(d_s, sdf + (size_t)o * BLOCK_VOX, (size_t)m * BLOCK_VOX * sizeof(float), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_w, weight + (size_t)o * BLOCK_VOX, (size_t)m * BLOCK_VOX * sizeof(float), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess && rgb) ce = cudaMemcpyAsync(d_c, rgb + (size_t)o * BLOCK_VOX * 3, (size_t)m * BLOCK_VOX * 3, cudaMemcpyHostToDevice, e->stream);
    upload_insert_kernel<<<(m + 255) / 256, 256, 0, e->stream>>>(e->D, e->d_keys_tmp, m, d_slots);
    upload_scatter_kernel<<<(m + 7) / 8, 256, 0, e->stream>>>(e->D, d_slots, m, d_s, d_w, rgb ? d_c : nullptr);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);      // `packed` and the caller's buffers are read by the copies above
    if (ce != cudaSuccess) rc = fail(VH_ERR_CUDA, "CUDA Error: %s in vh_upload_blocks", cudaGetErrorString(ce));
  }
  cudaFree(d_s); cudaFree(d_w); cudaFree(d_c); cudaFree(d_slots);
  if (rc != VH_OK) return rc;
  rc = enqueue_readback(e);                                            // heap counter, map error flags
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

}  // extern "C"
#endif  // !VH_HOST_EMU
