// vh_export.cu — read-back kernels: voxel planes of named blocks, checksums, per-block triangle records and the
// ordered triangle gather. These replace getMapValueKernel + streamOutGPU2CPU (/root/reference/src/tsdf.cu:459-596)
// and the dense slot scan of tsdf2mesh (tsdf.cu:1786-1806); they run on demand, never per frame.
#include "vh_engine.h"

namespace vh {

// one warp per requested block: lookup, then copy the 2 KB planes (float4 per lane x 4) into the staging arrays
__global__ void gather_blocks_kernel(const DeviceView D, const u64* __restrict__ keys, int n, float* __restrict__ sdf, float* __restrict__ wgt,
                                     uint8_t* __restrict__ rgb, uint8_t* __restrict__ found) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const int e = map_find(D.map, keys[w]);
  const int slot = e >= 0 ? D.map.slots[e] : -1;
  if (lane == 0) found[w] = slot >= 0;
  for (int v = lane; v < BLOCK_VOX; v += 32) {
    const size_t src = (size_t)(slot >= 0 ? slot : 0) * BLOCK_VOX + v, dst = (size_t)w * BLOCK_VOX + v;
    if (sdf) sdf[dst] = slot >= 0 ? D.sdf[src] : 0.0f;
    if (wgt) wgt[dst] = slot >= 0 ? D.wgt[src] : 0.0f;
    if (rgb) {
      uchar4 c = make_uchar4(0, 0, 0, 0);
      if (slot >= 0 && D.rgb) c = D.rgb[src];
      rgb[3 * dst] = c.x; rgb[3 * dst + 1] = c.y; rgb[3 * dst + 2] = c.z;
    }
  }
}

void launch_gather_blocks(const DeviceView& D, const u64* d_keys, int n, float* sdf, float* wgt, uint8_t* rgb, uint8_t* found, cudaStream_t st) {
  if (n <= 0) return;
  const int warps_per_cta = 8;
  gather_blocks_kernel<<<(n + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, st>>>(D, d_keys, n, sdf, wgt, rgb, found);
}

// out4 = {sum sdf, sum weight, #weight > 0, #sdf < 0} over every allocated block
__global__ void checksum_kernel(const DeviceView D, double* __restrict__ out4) {
  const int n = min(*D.map.heap_counter, D.map.num_blocks);
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  double ss = 0, sw = 0, no = 0, nn = 0;
  for (int i = w; i < n; i += nw) {
    const int e = map_find(D.map, D.map.key_heap[i]);
    if (e < 0) continue;
    const int slot = D.map.slots[e];
    if (slot < 0) continue;
    for (int v = lane; v < BLOCK_VOX; v += 32) {
      const float s = D.sdf[(size_t)slot * BLOCK_VOX + v], g = D.wgt[(size_t)slot * BLOCK_VOX + v];
      ss += s; sw += g; no += g > 0.0f; nn += s < 0.0f;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o); sw += __shfl_xor_sync(0xffffffffu, sw, o);
    no += __shfl_xor_sync(0xffffffffu, no, o); nn += __shfl_xor_sync(0xffffffffu, nn, o);
  }
  if (lane == 0) { atomicAdd(&out4[0], ss); atomicAdd(&out4[1], sw); atomicAdd(&out4[2], no); atomicAdd(&out4[3], nn); }
}

void launch_checksum(const DeviceView& D, double* d_out4, cudaStream_t st) {
  cudaMemsetAsync(d_out4, 0, 4 * sizeof(double), st);
  checksum_kernel<<<148 * 4, 256, 0, st>>>(D, d_out4);
}

// (key, arena offset, count) of every allocated block that holds triangles
__global__ void block_records_kernel(const DeviceView D, const unsigned long long* __restrict__ offsets, const int* __restrict__ counts,
                                     u64* __restrict__ rec_key, unsigned long long* __restrict__ rec_off, int* __restrict__ rec_cnt,
                                     int* __restrict__ n_out) {
  const int n = min(*D.map.heap_counter, D.map.num_blocks);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 key = D.map.key_heap[i];
  const int e = map_find(D.map, key);
  if (e < 0) return;
  const int slot = D.map.slots[e];
  if (slot < 0 || counts[slot] <= 0) return;
  const int pos = atomicAdd(n_out, 1);
  rec_key[pos] = key; rec_off[pos] = offsets[slot]; rec_cnt[pos] = counts[slot];
}

void launch_block_records(const DeviceView& D, const unsigned long long* offsets, const int* counts, u64* rec_key, unsigned long long* rec_off,
                          int* rec_cnt, int* n_out, cudaStream_t st) {
  cudaMemsetAsync(n_out, 0, sizeof(int), st);
  const int n = D.map.num_blocks;
  block_records_kernel<<<(n + 255) / 256, 256, 0, st>>>(D, offsets, counts, rec_key, rec_off, rec_cnt, n_out);
}

// one warp per block: copy its triangle range (48 B = 3 x uint4 each) to its place in the ordered output
__global__ void gather_triangles_kernel(const vh_triangle* __restrict__ arena, const unsigned long long* __restrict__ src_off,
                                        const unsigned long long* __restrict__ dst_off, const int* __restrict__ cnt, int n,
                                        vh_triangle* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const uint4* s = reinterpret_cast<const uint4*>(arena + src_off[w]);
  uint4* d = reinterpret_cast<uint4*>(out + dst_off[w]);
  const int q = cnt[w] * 3;
  for (int i = lane; i < q; i += 32) d[i] = s[i];
}

void launch_gather_triangles(const vh_triangle* arena, const unsigned long long* src_off, const unsigned long long* dst_off, const int* cnt, int n,
                             vh_triangle* out, cudaStream_t st) {
  if (n <= 0) return;
  gather_triangles_kernel<<<(n + 7) / 8, 256, 0, st>>>(arena, src_off, dst_off, cnt, n, out);
}

}  // namespace vh
