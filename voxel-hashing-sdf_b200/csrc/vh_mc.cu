// vh_mc.cu — marching cubes over a list of voxel blocks, two passes (count, emit) inside one CTA.
//
// Replaces marchingCubeHashKernel (/root/reference/src/tsdf.cu:884-1110): there every voxel thread performs
// 8 x (find + locked operator[]) hash walks to gather its cube corners and writes into a dense buffer of five
// 52-byte triangle slots per voxel (133 KB per 8^3 block) that is memset, filled and copied to the host every
// frame (tsdf.cu:1539-1565). Here one CTA owns one block:
//   1. 8 threads resolve the block and its 7 upper neighbours (one lock-free probe each) — a neighbour counts
//      only if it is in the same list (reference: "in this frame's working table", tsdf.cu:930,957-969);
//   2. the 9^3 sdf (+ colour) tile incl. the +x/+y/+z halo is staged in shared memory: each sdf is read from HBM once;
//   3. pass 1: cube index, edge vertices and the reference's degenerate-triangle rule give a per-voxel count;
//      a CTA scan turns counts into offsets and ONE atomicAdd reserves the block's range in the triangle arena;
//   4. pass 2: the valid triangles are written compactly (48 B each) in the reference's slot order (tid, k).
// The thread -> voxel mapping reproduces the reference's unsigned-arithmetic quirk (tsdf.cu:903-905);
// for 8^3 blocks it is the permutation y = ((tid >> 3) - bz) & 7, so slot order matches the reference's.
#include <cstdlib>

#include "vh_engine.h"
#include "vh_math.cuh"
#include "../../include/vh_mc_tables.h"

namespace vh {

constexpr int MC_THREADS = 512;
constexpr int TILE = 9, TILE_N = TILE * TILE * TILE;

// corner numbering of the reference's idxMap (tsdf.cuh:234-241) and Bourke's edge -> corner pairs, as arithmetic
// (same values as VH_MC_CORNER_OFFSET / VH_MC_EDGE_CORNERS in include/vh_mc_tables.h, usable in device code)
__host__ __device__ constexpr int corner_ox(int k) { return ((k & 3) >> 1); }
__host__ __device__ constexpr int corner_oy(int k) { return (((k & 3) == 1) || ((k & 3) == 2)) ? 1 : 0; }
__host__ __device__ constexpr int corner_oz(int k) { return k >> 2; }
__host__ __device__ constexpr int edge_a(int e) { return e < 8 ? e : e - 8; }
__host__ __device__ constexpr int edge_b(int e) { return e < 4 ? ((e + 1) & 3) : (e < 8 ? 4 + ((e - 3) & 3) : e - 4); }

__constant__ signed char c_tri[256][16];
__constant__ unsigned char c_ntri[256];
__constant__ unsigned short c_edge_mask[256];

void upload_mc_tables() {
  for (int k = 0; k < 8; k++)
    if (corner_ox(k) != VH_MC_CORNER_OFFSET[k][0] || corner_oy(k) != VH_MC_CORNER_OFFSET[k][1] || corner_oz(k) != VH_MC_CORNER_OFFSET[k][2]) abort();
  for (int e = 0; e < 12; e++)
    if (edge_a(e) != VH_MC_EDGE_CORNERS[e][0] || edge_b(e) != VH_MC_EDGE_CORNERS[e][1]) abort();
  static signed char tri[256][16];
  static unsigned char ntri[256];
  static unsigned short em[256];
  vh_mc_expand_tables(tri, ntri, em);
  cudaMemcpyToSymbol(c_tri, tri, sizeof(tri));
  cudaMemcpyToSymbol(c_ntri, ntri, sizeof(ntri));
  cudaMemcpyToSymbol(c_edge_mask, em, sizeof(em));
}

struct Vtx { float x, y, z; uint32_t c; };   // c = r | g<<8 | b<<16

// VertexInterp with isolevel 0 (tsdf.cu:1640-1660). The reference compares float fabs() results against the
// double literal 0.00001; the comparisons are done in double here for the same outcome.
__device__ __forceinline__ Vtx vertex_interp(const Vtx& p1, const Vtx& p2, float v1, float v2, bool color) {
  if ((double)fabsf(fsub(0.0f, v1)) < 0.00001) return p1;
  if ((double)fabsf(fsub(0.0f, v2)) < 0.00001) return p2;
  if ((double)fabsf(fsub(v1, v2)) < 0.00001) return p1;
  const float mu = fdiv(fsub(0.0f, v1), fsub(v2, v1));
  Vtx p;
  p.x = fadd(p1.x, fmul(mu, fsub(p2.x, p1.x)));
  p.y = fadd(p1.y, fmul(mu, fsub(p2.y, p1.y)));
  p.z = fadd(p1.z, fmul(mu, fsub(p2.z, p1.z)));
  p.c = 0;
  if (color) {
#pragma unroll
    for (int s = 0; s < 24; s += 8) {
      const int a = (p1.c >> s) & 0xFF, b = (p2.c >> s) & 0xFF;
      const int r = __float2int_rz(fadd(i2f(a), fmul(mu, i2f(b - a)))) & 0xFF;   // uchar arithmetic, truncating store
      p.c |= (uint32_t)r << s;
    }
  }
  return p;
}

__device__ __forceinline__ bool same_pos(const Vtx& a, const Vtx& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

__global__ void __launch_bounds__(MC_THREADS)
marching_cubes_kernel(const StaticParams S, const uint32_t frame, const DeviceView D, const int* __restrict__ list,
                      const int* __restrict__ list_count, const int full_map, unsigned long long* __restrict__ out_offset,
                      int* __restrict__ out_count) {
  __shared__ float s_sdf[TILE_N];
  __shared__ uint32_t s_rgb[TILE_N];
  __shared__ int s_nb_slot[8];
  __shared__ int s_warp_sum[MC_THREADS / 32];
  __shared__ unsigned long long s_base;
  __shared__ int s_total;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = min(*list_count, D.list_cap);
  const bool color = S.use_color != 0;
  unsigned long long my_tris = 0;

  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int entry = list[i];
    const u64 key = D.map.keys[entry];
    const int slot = D.map.slots[entry];
    int bx, by, bz;
    unpack_key(key, bx, by, bz);

    // 1. the block and its seven +x/+y/+z neighbours; bit0 = +x, bit1 = +y, bit2 = +z
    if (tid < 8) {
      int s = -1;
      if (tid == 0) s = slot;
      else {
        const int nx = bx + (tid & 1), ny = by + ((tid >> 1) & 1), nz = bz + ((tid >> 2) & 1);
        if (key_in_range(nx, ny, nz)) {
          const int e = map_find(D.map, pack_key(nx, ny, nz));
          if (e >= 0 && (full_map || D.stamps[e] == frame)) s = D.map.slots[e];
        }
      }
      s_nb_slot[tid] = s;
    }
    __syncthreads();

    // 2. 9^3 tile
    for (int c = tid; c < TILE_N; c += MC_THREADS) {
      const int tx = c / (TILE * TILE), ty = (c / TILE) % TILE, tz = c % TILE;
      const int nb = (tx == 8 ? 1 : 0) | (ty == 8 ? 2 : 0) | (tz == 8 ? 4 : 0);
      const int s = s_nb_slot[nb];
      float v = 0.0f;
      uint32_t col = 0;
      if (s >= 0) {
        const size_t a = (size_t)s * BLOCK_VOX + ((tx & 7) * 64 + (ty & 7) * 8 + (tz & 7));
        v = D.sdf[a];
        if (color) { const uchar4 q = D.rgb[a]; col = (uint32_t)q.x | ((uint32_t)q.y << 8) | ((uint32_t)q.z << 16); }
      }
      s_sdf[c] = v;
      s_rgb[c] = col;
    }
    __syncthreads();

    // 3. pass 1: per-voxel triangles. Reference thread mapping (tsdf.cu:903-906) for VPB = 8.
    const int lx = tid >> 6, ly = ((tid >> 3) - bz) & 7, lz = tid & 7;
    const int need = (lx == 7 ? 1 : 0) | (ly == 7 ? 2 : 0) | (lz == 7 ? 4 : 0);
    bool have = slot >= 0;
#pragma unroll
    for (int m = 1; m < 8; m++)
      if ((m & need) == m && s_nb_slot[m] < 0) have = false;     // a cube corner lies in a block that is not in the list

    int cube = 0;
    float val[8];
    if (have) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        val[k] = s_sdf[((lx + corner_ox(k)) * TILE + (ly + corner_oy(k))) * TILE + (lz + corner_oz(k))];
        cube |= (val[k] < 0.0f) ? (1 << k) : 0;                    // tsdf.cu:978-986 (no weight test, Q4)
      }
    }
    const int ntri = have ? c_ntri[cube] : 0;
    unsigned valid = 0;      // bit k: triangle k survives the degenerate rule
    Vtx tv[5][3];
    if (ntri) {
      const unsigned em = c_edge_mask[cube];
      Vtx vl[12];
#pragma unroll
      for (int e = 0; e < 12; e++) {
        if (em & (1u << e)) {
          const int a = edge_a(e), b = edge_b(e);
          Vtx pa, pb;
          const int ax = lx + corner_ox(a), ay = ly + corner_oy(a), az = lz + corner_oz(a);
          const int qx = lx + corner_ox(b), qy = ly + corner_oy(b), qz = lz + corner_oz(b);
          pa.x = i2f(bx * VPB + ax); pa.y = i2f(by * VPB + ay); pa.z = i2f(bz * VPB + az);          // Vertex(cxi, cyi, czi), tsdf.cu:931
          pb.x = i2f(bx * VPB + qx); pb.y = i2f(by * VPB + qy); pb.z = i2f(bz * VPB + qz);
          pa.c = s_rgb[(ax * TILE + ay) * TILE + az];
          pb.c = s_rgb[(qx * TILE + qy) * TILE + qz];
          vl[e] = vertex_interp(pa, pb, val[a], val[b], color);
        }
      }
      for (int k = 0; k < ntri; k++) {
        const Vtx p0 = vl[c_tri[cube][3 * k]], p1 = vl[c_tri[cube][3 * k + 1]], p2 = vl[c_tri[cube][3 * k + 2]];
        tv[k][0] = p0; tv[k][1] = p1; tv[k][2] = p2;
        if (!(same_pos(p0, p1) || same_pos(p1, p2))) valid |= 1u << k;   // p0 == p2 is never tested (Q5, tsdf.cu:1055-1057)
      }
    }
    const int cnt = __popc(valid);

    // CTA exclusive scan of cnt in tid order
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = lane < MC_THREADS / 32 ? s_warp_sum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      if (lane < MC_THREADS / 32) s_warp_sum[lane] = w;   // inclusive over warps
      if (lane == MC_THREADS / 32 - 1) {
        s_total = w;
        unsigned long long base = 0;
        if (w > 0) {
          base = atomicAdd(D.arena_top, (unsigned long long)w);
          if (base + (unsigned long long)w > D.arena_cap) { atomicOr(D.engine_error, 1); s_total = -w; }
        }
        s_base = base;
      }
    }
    __syncthreads();
    const int total = s_total;
    const unsigned long long base = s_base;
    if (tid == 0 && slot >= 0) {
      out_offset[slot] = base;
      out_count[slot] = total > 0 ? total : 0;
    }
    // 4. pass 2: emit
    if (cnt && total > 0) {
      unsigned long long pos = base + (unsigned long long)((wid ? s_warp_sum[wid - 1] : 0) + incl - cnt);
      for (int k = 0; k < ntri; k++) {
        if (valid & (1u << k)) {
          uint4* dst = reinterpret_cast<uint4*>(D.arena + pos);
#pragma unroll
          for (int j = 0; j < 3; j++)
            dst[j] = make_uint4(__float_as_uint(tv[k][j].x), __float_as_uint(tv[k][j].y), __float_as_uint(tv[k][j].z), tv[k][j].c);
          ++pos;
        }
      }
    }
    if (tid == 0 && total > 0) my_tris += (unsigned long long)total;
    __syncthreads();   // smem reuse by the next block
  }
  if (tid == 0 && my_tris && !full_map) atomicAdd(&D.counters->triangles, my_tris);
}

void launch_marching_cubes(const StaticParams& S, const FrameParams& F, const DeviceView& D, const int* list, const int* list_count, int full_map,
                           unsigned long long* out_offset, int* out_count, int num_sms, cudaStream_t st) {
  marching_cubes_kernel<<<num_sms * 4, MC_THREADS, 0, st>>>(S, F.frame, D, list, list_count, full_map, out_offset, out_count);
}

}  // namespace vh
