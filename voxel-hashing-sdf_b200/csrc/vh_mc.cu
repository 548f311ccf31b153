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

static uint4* g_tables_dev[16] = {};          // per device: expanded case tables in global memory (mesh kernel copies them to shared)

#ifndef VH_HOST_EMU      // (the CPU emulation of the kernels, tests/emu, expands the tables itself)
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
  int dev = 0;
  cudaGetDevice(&dev);
  if (!g_tables_dev[dev & 15]) {
    cudaMalloc((void**)&g_tables_dev[dev & 15], sizeof(tri) + sizeof(ntri));
    cudaMemcpy(g_tables_dev[dev & 15], tri, sizeof(tri), cudaMemcpyHostToDevice);
    cudaMemcpy(reinterpret_cast<char*>(g_tables_dev[dev & 15]) + sizeof(tri), ntri, sizeof(ntri), cudaMemcpyHostToDevice);
  }
}

#endif  // !VH_HOST_EMU

struct Vtx { float x, y, z; uint32_t c; };   // c = r | g<<8 | b<<16

// VertexInterp with isolevel 0 (tsdf.cu:1640-1660). The reference compares float fabs() results against the
// double literal 0.00001; the equivalent float threshold is used here.
__device__ __forceinline__ Vtx vertex_interp(const Vtx& p1, const Vtx& p2, float v1, float v2, bool color) {
  // (double)|x| < 0.00001  <=>  |x| <= 0x3727C5AC (the largest float below the double 1e-5; its successor is above it)
  const float eps = __uint_as_float(0x3727C5ACu);
  if (fabsf(fsub(0.0f, v1)) <= eps) return p1;
  if (fabsf(fsub(0.0f, v2)) <= eps) return p2;
  if (fabsf(fsub(v1, v2)) <= eps) return p1;
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

constexpr int MC_WARPS = 4;                 // voxel blocks in flight per CTA (one warp each); 36.7 KB of static shared memory
constexpr int MC_THREADS = MC_WARPS * 32;
constexpr int TILE = 9, TILE_PAD = 736;   // 9^3 = 729 cells, padded

struct McBlock {            // per-warp context of the block being meshed
  const float* tile;        // 9^3 sdf tile in shared memory
  const uint32_t* ctile;    // colour tile in shared memory (CTILE kernels): [512] the block's own voxels in plane order, [512 + c] halo cell c
  const int* nb_slot;       // pool slots of the 8 corner blocks (bit0 +x, bit1 +y, bit2 +z), -1 = absent
  const int* nb_owner;      // multi-GPU: shard that holds each corner block (its planes are read through D.peers)
  int bx, by, bz;
};

// The vertex on cube edge e of the voxel at tile position (lx,ly,lz), in two halves: edge_fetch (addresses, the two sdf values,
// the two colour gathers) and edge_finish (positions, VertexInterp). A triangle's three fetches are issued before the first
// finish, so its six colour gathers are all in flight before the first interpolation consumes one (measured on B200 against
// fetch-and-interpolate one vertex at a time: 0.069 -> 0.066 ms per frame, profiles/r02a).
// asynchronous global -> shared copies without registers (LDGSTS): the colour tile travels while pass 1 runs
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
#ifdef VH_HOST_EMU
  memcpy(dst_smem, src, 16);
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src) {
#ifdef VH_HOST_EMU
  memcpy(dst_smem, src, 4);
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef VH_HOST_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef VH_HOST_EMU
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// colour of the tile cell (x, y, z), 0..8 each, out of the staged colour tile
__device__ __forceinline__ uint32_t ctile_color(const uint32_t* ctile, int x, int y, int z) {
  int i;
  if (((x | y | z) & 8) == 0) i = x * 64 + y * 8 + z;
  else if (x == 8) i = 512 + y * 9 + z;
  else if (y == 8) i = 512 + 81 + x * 9 + z;
  else i = 512 + 153 + x * 8 + y;
  return ctile[i] & 0xFFFFFFu;
}

struct EdgeFetch { int a_pos, b_pos; float va, vb; uint32_t ca, cb; };   // positions packed x | y << 8 | z << 16 (tile coordinates 0..8)
template <bool SHARDED, bool CTILE>
__device__ __forceinline__ EdgeFetch edge_fetch(const McBlock& B, const DeviceView& D, int lx, int ly, int lz, int e, bool color) {
  const int a = e < 8 ? e : e - 8;
  const int b = e < 4 ? ((e + 1) & 3) : (e < 8 ? 4 + ((e - 3) & 3) : e - 4);
  const int ax = lx + corner_ox(a), ay = ly + corner_oy(a), az = lz + corner_oz(a);
  const int qx = lx + corner_ox(b), qy = ly + corner_oy(b), qz = lz + corner_oz(b);
  EdgeFetch f;
  f.a_pos = ax | (ay << 8) | (az << 16); f.b_pos = qx | (qy << 8) | (qz << 16);
  f.ca = 0; f.cb = 0;
  if (color && CTILE) {
    f.ca = ctile_color(B.ctile, ax, ay, az); f.cb = ctile_color(B.ctile, qx, qy, qz);
  } else if (color) {
    const int ma = (ax >> 3) | ((ay >> 3) << 1) | ((az >> 3) << 2), mb = (qx >> 3) | ((qy >> 3) << 1) | ((qz >> 3) << 2);
    const int sa = B.nb_slot[ma], sb = B.nb_slot[mb];
    const uchar4* rgb_a = SHARDED ? D.peers->v[B.nb_owner[ma]].rgb : D.rgb;
    const uchar4* rgb_b = SHARDED ? D.peers->v[B.nb_owner[mb]].rgb : D.rgb;
    const uchar4 ca = rgb_a[(size_t)sa * BLOCK_VOX + ((ax & 7) * 64 + (ay & 7) * 8 + (az & 7))];
    const uchar4 cb = rgb_b[(size_t)sb * BLOCK_VOX + ((qx & 7) * 64 + (qy & 7) * 8 + (qz & 7))];
    f.ca = (uint32_t)ca.x | ((uint32_t)ca.y << 8) | ((uint32_t)ca.z << 16);
    f.cb = (uint32_t)cb.x | ((uint32_t)cb.y << 8) | ((uint32_t)cb.z << 16);
  }
  f.va = B.tile[(ax * TILE + ay) * TILE + az]; f.vb = B.tile[(qx * TILE + qy) * TILE + qz];
  return f;
}
__device__ __forceinline__ Vtx edge_finish(const McBlock& B, const EdgeFetch& f, bool color) {
  Vtx pa, pb;
  pa.x = i2f(B.bx * VPB + (f.a_pos & 0xFF)); pa.y = i2f(B.by * VPB + ((f.a_pos >> 8) & 0xFF)); pa.z = i2f(B.bz * VPB + (f.a_pos >> 16));
  pb.x = i2f(B.bx * VPB + (f.b_pos & 0xFF)); pb.y = i2f(B.by * VPB + ((f.b_pos >> 8) & 0xFF)); pb.z = i2f(B.bz * VPB + (f.b_pos >> 16));
  pa.c = f.ca; pb.c = f.cb;
  return vertex_interp(pa, pb, f.va, f.vb, color);
}

__device__ __forceinline__ int cube_index(const float* tile, int lx, int ly, int lz) {
  int cube = 0;
#pragma unroll
  for (int k = 0; k < 8; k++)
    cube |= (tile[((lx + corner_ox(k)) * TILE + (ly + corner_oy(k))) * TILE + (lz + corner_oz(k))] < 0.0f) ? (1 << k) : 0;   // tsdf.cu:978-986, no weight test (Q4)
  return cube;
}

// Mesh one block with the whole warp: stage the 9^3 tile, pass 1 lists candidate triangles, pass 2 writes the survivors.
// B.nb_slot[0..8) holds the pool slots of the block and its seven upper neighbours (-1 = absent), `present` the same as bits.
// cube_cache keeps the cube index of every voxel that has triangles (one byte per voxel per warp) so that the emit pass does not
// rebuild it from eight shared-memory reads per candidate triangle.
template <bool SHARDED, bool CTILE>
__device__ __forceinline__ int mesh_block(const McBlock& B, const DeviceView& D, const int slot, const unsigned present, const int (&halo)[7],
                                          float* tile, uint32_t* ctile, unsigned short* wlist, unsigned char* cube_cache, const signed char* s_tri, const unsigned char* s_ntri,
                                          const bool color, unsigned long long* __restrict__ out_offset, int* __restrict__ out_count,
                                          const int lane, const uint32_t frame) {
    // okbits bit q: every corner block a voxel with boundary mask q touches is present
    bool okq = false;
    if (lane < 8) {
      okq = true;
#pragma unroll
      for (int m = 0; m < 8; m++) if ((m & lane) == m && !((present >> m) & 1u)) okq = false;
    }
    const unsigned okbits = __ballot_sync(0xffffffffu, okq) & 0xFFu;

    // 9^3 tile: own 8^3 as four float4 per lane, then the 217 halo cells. While loading, note whether any cell is
    // negative / non-negative: a tile of one sign class has cube index 0 or 255 everywhere and yields no triangle.
    bool any_neg = false, any_pos = false;
    {
      const float* src = D.sdf + (size_t)slot * BLOCK_VOX;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int v = (j * 32 + lane) * 4;
        const float4 q = *reinterpret_cast<const float4*>(src + v);
        float* dst = tile + ((v >> 6) * TILE + ((v >> 3) & 7)) * TILE + (v & 7);
        dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
        any_neg = any_neg || q.x < 0.0f || q.y < 0.0f || q.z < 0.0f || q.w < 0.0f;
        any_pos = any_pos || !(q.x < 0.0f) || !(q.y < 0.0f) || !(q.z < 0.0f) || !(q.w < 0.0f);
      }
    }
#pragma unroll
    for (int h = 0; h < 7; h++) {
      const int d = halo[h];
      if (d >= 0) {
        const int s = B.nb_slot[(d >> 9) & 7];
        const float* plane = SHARDED ? D.peers->v[B.nb_owner[(d >> 9) & 7]].sdf : D.sdf;
        const float v = s >= 0 ? plane[(size_t)s * BLOCK_VOX + (d & 511)] : 0.0f;
        tile[d >> 12] = v;
        any_neg = any_neg || v < 0.0f;
        any_pos = any_pos || !(v < 0.0f);
      }
    }
    const bool mixed = __any_sync(0xffffffffu, any_neg) && __any_sync(0xffffffffu, any_pos);
    if (!mixed) {
      if (lane == 0) { out_offset[slot] = 0; out_count[slot] = 0; }
      return 0;
    }
    __syncwarp();
    // CTILE: the block's colours (2 KB, coalesced) and the 217 halo colours start travelling into shared memory now, without
    // registers, and arrive while pass 1 runs; pass 2 then interpolates colours out of shared memory instead of six dependent
    // global gathers per triangle (the mesh kernel's top stall, profiles/r02_ncu/ncu_source_mcmesh.txt)
    if (CTILE && color) {
      const uchar4* own = (SHARDED ? D.peers->v[B.nb_owner[0]].rgb : D.rgb) + (size_t)slot * BLOCK_VOX;
#pragma unroll
      for (int j = 0; j < 4; j++) cp_async_16(ctile + (j * 32 + lane) * 4, own + (j * 32 + lane) * 4);
#pragma unroll
      for (int h = 0; h < 7; h++) {
        const int d = halo[h];
        if (d >= 0) {
          const int s = B.nb_slot[(d >> 9) & 7];
          const uchar4* plane = SHARDED ? D.peers->v[B.nb_owner[(d >> 9) & 7]].rgb : D.rgb;
          if (s >= 0) cp_async_4(ctile + 512 + h * 32 + lane, plane + (size_t)s * BLOCK_VOX + (d & 511));
          else ctile[512 + h * 32 + lane] = 0u;
        }
      }
      cp_async_commit();
    }

    // pass 1 (count): list the candidate triangles of the block as (tid << 3 | k) in the reference's slot order.
    // Reference thread -> voxel mapping (tsdf.cu:903-906) for VPB = 8: tid = j*32 + lane -> x = tid >> 6,
    // y = ((tid >> 3) - bz) & 7, z = tid & 7.
    unsigned short* wl = wlist;
    int nlist = 0;
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
      const int t = j * 32 + lane;
      const int lx = t >> 6, ly = ((t >> 3) - B.bz) & 7, lz = t & 7;
      const int need = (lx == 7 ? 1 : 0) | (ly == 7 ? 2 : 0) | (lz == 7 ? 4 : 0);
      int nt = 0;
      if ((okbits >> need) & 1u) {
        const int ci = cube_index(tile, lx, ly, lz);
        nt = s_ntri[ci];                                                         // 0 for cube index 0 and 255
        if (nt > 0) cube_cache[t] = (unsigned char)ci;
      }
      if (!__any_sync(0xffffffffu, nt > 0)) continue;
      int incl = nt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
      for (int k = 0; k < nt; k++) wl[nlist + incl - nt + k] = (unsigned short)((t << 3) | k);
      nlist += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();

    // one reservation per block for all candidates (the few degenerate ones leave unused arena slots behind the block's range)
    unsigned long long base = 0;
    bool fits = true;
    if (nlist > 0) {
      if (lane == 0) {
        base = atomicAdd(D.arena_top, (unsigned long long)nlist);
        if (base + (unsigned long long)nlist > D.arena_cap) { atomicOr(D.engine_error, 1); atomicCAS(D.overflow_frame, 0u, frame); fits = false; }
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      fits = __shfl_sync(0xffffffffu, fits, 0);
    }

    if (CTILE && color) { cp_async_wait_all(); __syncwarp(); }
    // pass 2 (emit): one candidate triangle per lane; survivors of the degenerate rule are written compactly, in order
    int written = 0;
    if (nlist > 0 && fits) {
      for (int e0 = 0; e0 < nlist; e0 += 32) {
        const int e = e0 + lane;
        bool valid = false;
        Vtx p0, p1, p2;
        if (e < nlist) {
          const int item = wl[e];
          const int t = item >> 3, k = item & 7;
          const int lx = t >> 6, ly = ((t >> 3) - B.bz) & 7, lz = t & 7;
          const signed char* row = s_tri + (int)cube_cache[t] * 16 + 3 * k;
          const EdgeFetch f0 = edge_fetch<SHARDED, CTILE>(B, D, lx, ly, lz, row[0], color), f1 = edge_fetch<SHARDED, CTILE>(B, D, lx, ly, lz, row[1], color),
                          f2 = edge_fetch<SHARDED, CTILE>(B, D, lx, ly, lz, row[2], color);
          p0 = edge_finish(B, f0, color); p1 = edge_finish(B, f1, color); p2 = edge_finish(B, f2, color);
          valid = !(same_pos(p0, p1) || same_pos(p1, p2));                       // p0 == p2 is never tested (Q5, tsdf.cu:1055-1057)
        }
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(D.arena + base + (unsigned long long)(written + __popc(bal & ((1u << lane) - 1))));
          dst[0] = make_uint4(__float_as_uint(p0.x), __float_as_uint(p0.y), __float_as_uint(p0.z), p0.c);
          dst[1] = make_uint4(__float_as_uint(p1.x), __float_as_uint(p1.y), __float_as_uint(p1.z), p1.c);
          dst[2] = make_uint4(__float_as_uint(p2.x), __float_as_uint(p2.y), __float_as_uint(p2.z), p2.c);
        }
        written += __popc(bal);
      }
    }
    if (lane == 0) { out_offset[slot] = base; out_count[slot] = written; }
    return written;
}

// ---- stage 1: filter ------------------------------------------------------------------------------------------------
// The list is consumed four blocks at a time per warp: lane group g = lane >> 3 resolves block base+g and its seven
// +x/+y/+z neighbours (one lock-free probe per lane) and reads their negative-voxel counters (maintained by the
// integrate kernel). A block goes on to meshing only if, among the blocks its 9^3 tile draws from, some voxel is
// negative and some is not — otherwise every cube index is 0 or 255 and the reference emits nothing for it either. In a
// room scan most of the working set is free space, so most blocks end here without their voxels being read. Survivors
// are appended to a work queue (block, slot, the eight corner-block slots and owners), which decouples meshing from the
// list order: surface blocks are clustered in the list, and a warp that drew four of them used to serialise them.
// The neighbour look-up reads key, stamp and slot of the FIRST probe position together (three independent loads; at the
// table's load factor nine look-ups in ten end there) instead of key -> stamp/slot one after the other, which shortens the
// chain of dependent loads the filter spends its time waiting on. Later probe positions (rare) take the ordinary path.
template <bool SHARDED>
__global__ void __launch_bounds__(256)
mc_filter_kernel(const StaticParams S, const uint32_t frame, const DeviceView D, const int* __restrict__ list,
                 const int* __restrict__ list_count, const int full_map, unsigned long long* __restrict__ out_offset,
                 int* __restrict__ out_count, McWork* __restrict__ queue, McQueueCtl* __restrict__ ctl) {
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n = min(*list_count, D.list_cap);
  for (int base = gwarp * 4; base < n; base += nwarps * 4) {
    const int grp = lane >> 3, sub = lane & 7;
    const int item = base + grp;
    int slot = -1, bx = 0, by = 0, bz = 0;
    if (item < n) {
      const int entry = list[item];
      const u64 key = D.map.keys[entry];
      slot = D.map.slots[entry];
      unpack_key(key, bx, by, bz);
    }
    // a neighbour counts only if it is in the same list (this frame's working set, tsdf.cu:930,957-969) or, for
    // full-map extraction, allocated at all
    int nb = -1, nb_owner = SHARDED ? (int)S.shard_rank : 0, nneg = 0;
    if (sub == 0) { nb = slot; if (slot >= 0) nneg = D.neg_count[slot]; }
    else if (slot >= 0) {
      const int nx = bx + (sub & 1), ny = by + ((sub >> 1) & 1), nz = bz + ((sub >> 2) & 1);
      if (key_in_range(nx, ny, nz)) {
        const u64 nk = pack_key(nx, ny, nz);
        const u64* t_keys = D.map.keys; const int* t_slots = D.map.slots; const uint32_t* t_stamps = D.stamps; const int* t_neg = D.neg_count;
        uint32_t t_mask = D.map.mask;
        if (SHARDED) {
          // the neighbour lives on the GPU its key hashes to: probe that GPU's table and read its stamp / counter over NVLink
          nb_owner = (int)owner_of_block(nx, ny, nz, S.shard_count, S.shard_group);
          const PeerView P = D.peers->v[nb_owner];
          t_keys = P.keys; t_slots = P.slots; t_stamps = P.stamps; t_neg = P.neg_count; t_mask = P.mask;
        }
        const uint32_t h0 = hash_key(nk) & t_mask;
        const u64 k0 = __ldcg(&t_keys[h0]);
        uint32_t st = t_stamps[h0];
        int sl = t_slots[h0];
        int e = k0 == nk ? (int)h0 : -1;
        if (k0 != nk && k0 != KEY_EMPTY) { e = map_find_in(t_keys, t_mask, nk); if (e >= 0) { st = t_stamps[e]; sl = t_slots[e]; } }
        if (e >= 0 && (full_map || st == frame)) { nb = sl; if (nb >= 0) nneg = t_neg[nb]; }
      }
    }
    const unsigned gsh = grp * 8;
    const unsigned b_present = __ballot_sync(0xffffffffu, nb >= 0);
    const unsigned b_neg = __ballot_sync(0xffffffffu, nb >= 0 && nneg > 0);
    const unsigned b_pos = __ballot_sync(0xffffffffu, nb >= 0 && nneg < BLOCK_VOX);
    const bool need = slot >= 0 && ((b_neg >> gsh) & 0xFFu) != 0 && ((b_pos >> gsh) & 0xFFu) != 0;
    if (sub == 0 && slot >= 0 && !need) { out_offset[slot] = 0; out_count[slot] = 0; }
    // one queue reservation per warp
    const unsigned lead = __ballot_sync(0xffffffffu, need && sub == 0);
    if (lead) {
      int qbase = 0;
      if (lane == 0) qbase = atomicAdd(&ctl->count, __popc(lead));
      qbase = __shfl_sync(0xffffffffu, qbase, 0);
      if (need) {
        McWork* w = queue + qbase + __popc(lead & ((1u << (grp * 8)) - 1));
        w->nb[sub] = nb;
        if (SHARDED) w->owner[sub] = (unsigned char)nb_owner;
        if (sub == 0) { w->bx = bx; w->by = by; w->bz = bz; w->slot = slot; w->present = (b_present >> gsh) & 0xFFu; }
      }
    }
  }
}

// ---- stage 2: mesh ----------------------------------------------------------------------------------------------------
// Persistent warps pull blocks off the work queue (one atomicAdd per block) and mesh them: perfect balance whatever the
// spatial clustering of surface blocks. The last thing the kernel does is clear the OTHER queue-control slot, which the
// next launch pair will use (the two slots alternate, so no memset node is needed per frame).
template <bool SHARDED, bool CTILE>
__global__ void __launch_bounds__(MC_THREADS)
mc_mesh_kernel(const StaticParams S, const uint32_t frame, const DeviceView D, const int full_map, unsigned long long* __restrict__ out_offset,
               int* __restrict__ out_count, const McWork* __restrict__ queue, McQueueCtl* __restrict__ ctl, McQueueCtl* __restrict__ ctl_next,
               const uint4* __restrict__ tables) {
  __shared__ float s_tile[MC_WARPS][TILE_PAD];
  __shared__ int s_nb[MC_WARPS][8];
  __shared__ int s_nbo[MC_WARPS][8];
  __shared__ unsigned short s_list[MC_WARPS][BLOCK_VOX * 5];   // candidate triangles of the block in flight
  __shared__ __align__(16) signed char s_tri[256 * 16];
  __shared__ __align__(16) unsigned char s_ntri[256];
  __shared__ unsigned char s_cube[MC_WARPS][BLOCK_VOX];        // cube index of the voxels that have triangles
#ifdef VH_HOST_EMU
  uint32_t* dyn_ctile = reinterpret_cast<uint32_t*>(emu::g_cta->dyn_smem);
#else
  extern __shared__ __align__(16) uint32_t dyn_ctile[];        // CTILE: [MC_WARPS][TILE_PAD] colour tiles (dynamic: with it the CTA holds more than 48 KB)
#endif

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int total = ctl->count;
  if (blockIdx.x == 0 && tid == 0) { ctl_next->count = 0; ctl_next->head = 0; }
  if ((int)blockIdx.x * MC_WARPS >= total) return;                 // nothing left for this CTA: skip the table copy too
  // case tables: 4 KB + 256 B, copied as 16-byte words from a global copy (constant-bank byte loads serialise)
  for (int i = tid; i < 256; i += MC_THREADS) reinterpret_cast<uint4*>(s_tri)[i] = tables[i];
  for (int i = tid; i < 16; i += MC_THREADS) reinterpret_cast<uint4*>(s_ntri)[i] = tables[256 + i];
  __syncthreads();

  const bool color = S.use_color != 0;
  float* tile = s_tile[wid];
  unsigned long long my_tris = 0;

  // halo cells handled by this lane, the same for every block: tile offset << 12 | corner-block index << 9 | voxel in that block
  int halo[7];
#pragma unroll
  for (int h = 0; h < 7; h++) {
    const int c = h * 32 + lane;
    int tx, ty, tz;
    if (c < 81) { tx = 8; ty = c / 9; tz = c - ty * 9; }
    else if (c < 153) { const int q = c - 81; tx = q / 9; ty = 8; tz = q - tx * 9; }
    else { const int q = c - 153; tx = q >> 3; ty = q & 7; tz = 8; }
    const int m = (tx >> 3) | ((ty >> 3) << 1) | ((tz >> 3) << 2);
    halo[h] = c < 217 ? ((((tx * TILE + ty) * TILE + tz) << 12) | (m << 9) | ((tx & 7) * 64 + (ty & 7) * 8 + (tz & 7))) : -1;
  }

  // (Popping the next queue item ahead of the block in flight was measured and dropped: with ~2 blocks per warp a warp that holds
  //  its next item while it meshes keeps it from an idle warp — 0.059 -> 0.073 ms per frame on config 2, profiles/r02final.)
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&ctl->head, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= total) break;
    const McWork* w = queue + idx;
    McBlock C;
    C.bx = w->bx; C.by = w->by; C.bz = w->bz; C.tile = tile; C.ctile = CTILE ? dyn_ctile + wid * TILE_PAD : nullptr; C.nb_slot = s_nb[wid]; C.nb_owner = s_nbo[wid];
    const int cur_slot = w->slot;
    const unsigned present = w->present;
    __syncwarp();                              // previous block's readers of s_nb / tile / colour tile / list are done
    if (lane < 8) { s_nb[wid][lane] = w->nb[lane]; s_nbo[wid][lane] = SHARDED ? (int)w->owner[lane] : 0; }
    __syncwarp();
    my_tris += (unsigned long long)mesh_block<SHARDED, CTILE>(C, D, cur_slot, present, halo, tile, CTILE ? dyn_ctile + wid * TILE_PAD : nullptr, s_list[wid], s_cube[wid], s_tri, s_ntri, color, out_offset, out_count, lane, frame);
  }
  if (lane == 0 && my_tris && !full_map) atomicAdd(&D.counters->triangles, my_tris);
}

#ifndef VH_HOST_EMU
void launch_marching_cubes(const StaticParams& S, const FrameParams& F, const DeviceView& D, const int* list, const int* list_count, int full_map,
                           unsigned long long* out_offset, int* out_count, int num_sms, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  McQueueCtl* ctl = D.mc_ctl + (*D.mc_parity & 1);
  McQueueCtl* ctl_next = D.mc_ctl + ((*D.mc_parity & 1) ^ 1);
  *D.mc_parity ^= 1;
  // launch shapes; the environment overrides exist for sweeps on the GPU box (both kernels are latency-bound: the filter
  // walks chains of dependent look-ups at 26 registers per thread, so 8 CTAs per SM fit: 0.069 -> 0.063 ms per frame against 4, profiles/r02a)
  static const int filter_ctas = [] { const char* v = getenv("VH_MC_FILTER_CTAS"); const int n = v ? atoi(v) : 0; return n >= 1 && n <= 8 ? n : 8; }();
  static const int mesh_ctas = [] { const char* v = getenv("VH_MC_MESH_CTAS"); const int n = v ? atoi(v) : 0; return n >= 1 && n <= 6 ? n : 5; }();
  const int fgrid = num_sms * filter_ctas;
  const int mgrid = num_sms * mesh_ctas;         // 5 CTAs of 4 warps fit an SM (36.7 KB of shared memory each)
  // colour tile through shared memory (VH_MC_COLOR_TILE, see mesh_block): 11.5 KB of dynamic shared memory more, 4 instead of 5 CTAs per SM
  const bool use_ctile = S.mc_rev == 1 && S.use_color;
  const size_t dyn = use_ctile ? (size_t)MC_WARPS * TILE_PAD * sizeof(uint32_t) : 0;
  const bool sharded = S.shard_count > 1 && D.peers;
#define VH_MESH(SH, CT) do { \
    if (CT) { static bool done[16] = {}; if (!done[dev & 15]) { cudaFuncSetAttribute(mc_mesh_kernel<SH, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn); done[dev & 15] = true; } } \
    mc_mesh_kernel<SH, CT><<<(CT ? num_sms * 4 : mgrid), MC_THREADS, dyn, st>>>(S, F.frame, D, full_map, out_offset, out_count, D.mc_queue, ctl, ctl_next, g_tables_dev[dev & 15]); } while (0)
  if (sharded) {
    mc_filter_kernel<true><<<fgrid, 256, 0, st>>>(S, F.frame, D, list, list_count, full_map, out_offset, out_count, D.mc_queue, ctl);
    if (use_ctile) VH_MESH(true, true); else VH_MESH(true, false);
  } else {
    mc_filter_kernel<false><<<fgrid, 256, 0, st>>>(S, F.frame, D, list, list_count, full_map, out_offset, out_count, D.mc_queue, ctl);
    if (use_ctile) VH_MESH(false, true); else VH_MESH(false, false);
  }
#undef VH_MESH
}
#endif  // !VH_HOST_EMU

}  // namespace vh
