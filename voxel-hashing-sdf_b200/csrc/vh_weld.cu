// vh_weld.cu — vertex welding of the ordered triangle soup on the GPU.
//
// Replaces the vertex dedupe of tsdf2mesh (/root/reference/src/tsdf.cu:1777,1810-1821: a std::unordered_map over every
// triangle vertex keyed on exact float xyz, first occurrence's colour wins, ids in order of first appearance, coordinates
// multiplied by vox_size). Same result, computed with two stable radix sorts instead of a host hash map:
//   1. key every soup vertex i by its (x, y, z) bits (-0 folded onto +0, as float == does) and sort the positions by z,
//      then stably by (x, y): equal vertices become adjacent, in ascending soup position;
//   2. the first position of each run is the vertex's first occurrence = its representative;
//   3. representatives, taken in soup order, get consecutive ids (exclusive scan) and emit the welded vertex;
//      every soup vertex maps to its representative's id = the face index list.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "vh_engine_host.h"

namespace {

__device__ __forceinline__ const vh_vertex& soup_vertex(const vh_triangle* __restrict__ soup, unsigned i) { return soup[i / 3].p[i % 3]; }

__global__ void weld_keys_kernel(const vh_triangle* __restrict__ soup, unsigned n, u64* __restrict__ kxy, unsigned* __restrict__ kz, unsigned* __restrict__ pos) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const vh_vertex& v = soup_vertex(soup, i);
  const float x = v.x + 0.0f, y = v.y + 0.0f, z = v.z + 0.0f;        // -0 -> +0
  kxy[i] = ((u64)__float_as_uint(x) << 32) | (u64)__float_as_uint(y);
  kz[i] = __float_as_uint(z);
  pos[i] = i;
}
__global__ void weld_gather_kernel(const u64* __restrict__ kxy, const unsigned* __restrict__ pos, unsigned n, u64* __restrict__ out) {
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = kxy[pos[j]];
}
__global__ void weld_heads_kernel(const u64* __restrict__ kxy, const unsigned* __restrict__ kz, const unsigned* __restrict__ sorted_pos, unsigned n,
                                  unsigned* __restrict__ head_index) {
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  bool head = j == 0;
  if (!head) { const unsigned a = sorted_pos[j], b = sorted_pos[j - 1]; head = kxy[a] != kxy[b] || kz[a] != kz[b]; }
  head_index[j] = head ? j : 0u;
}
__global__ void weld_reps_kernel(const unsigned* __restrict__ sorted_pos, const unsigned* __restrict__ run_head, unsigned n, unsigned* __restrict__ rep,
                                 unsigned* __restrict__ is_rep) {
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned i = sorted_pos[j], r = sorted_pos[run_head[j]];     // positions ascend inside a run: the head is the first occurrence
  rep[i] = r;
  is_rep[i] = i == r ? 1u : 0u;
}
__global__ void weld_emit_kernel(const vh_triangle* __restrict__ soup, unsigned n, const unsigned* __restrict__ rep, const unsigned* __restrict__ is_rep,
                                 const unsigned* __restrict__ id, float vox_size, vh_vertex* __restrict__ verts, int32_t* __restrict__ faces) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  faces[i] = (int32_t)id[rep[i]];
  if (is_rep[i]) {
    vh_vertex v = soup_vertex(soup, i);
    v.x = __fmul_rn(v.x, vox_size); v.y = __fmul_rn(v.y, vox_size); v.z = __fmul_rn(v.z, vox_size); v.pad = 0;   // tsdf.cu:1815-1817
    verts[id[i]] = v;
  }
}

struct MaxOp { __host__ __device__ unsigned operator()(unsigned a, unsigned b) const { return a > b ? a : b; } };

struct DevBuf {       // frees on scope exit
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

}  // namespace

// d_soup: T triangles in mesh order on the device. Fills the welded vertices (world units) and the face index list.
int weld_on_device(vh_engine* e, const vh_triangle* d_soup, unsigned long long T, std::vector<vh_vertex>& verts, std::vector<int32_t>& faces) {
  verts.clear(); faces.clear();
  if (T == 0) return VH_OK;
  if (T * 3ull >= (1ull << 31)) return fail(VH_ERR_INVALID, "mesh of %llu triangles is too large to weld in one pass", T);
  const unsigned n = (unsigned)(T * 3ull);
  cudaStream_t st = e->stream;
  DevBuf kxy, kz, pos, kxy_g, kxy_s, kz_s, pos_s, pos2, head, run, rep, isrep, id, tmp, dverts, dfaces;
#define WALLOC(buf, bytes) CK(cudaMalloc(&buf.p, (bytes)))
  WALLOC(kxy, (size_t)n * 8); WALLOC(kz, (size_t)n * 4); WALLOC(pos, (size_t)n * 4);
  WALLOC(kz_s, (size_t)n * 4); WALLOC(pos_s, (size_t)n * 4);
  WALLOC(kxy_g, (size_t)n * 8); WALLOC(kxy_s, (size_t)n * 8); WALLOC(pos2, (size_t)n * 4);
  const unsigned g = (n + 255) / 256;
  weld_keys_kernel<<<g, 256, 0, st>>>(d_soup, n, kxy.as<u64>(), kz.as<unsigned>(), pos.as<unsigned>());
  size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, kz.as<unsigned>(), kz_s.as<unsigned>(), pos.as<unsigned>(), pos_s.as<unsigned>(), (int)n, 0, 32, st);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, kxy_g.as<u64>(), kxy_s.as<u64>(), pos_s.as<unsigned>(), pos2.as<unsigned>(), (int)n, 0, 64, st);
  cub::DeviceScan::InclusiveScan(nullptr, b3, (unsigned*)nullptr, (unsigned*)nullptr, MaxOp(), (int)n, st);
  cub::DeviceScan::ExclusiveSum(nullptr, b4, (unsigned*)nullptr, (unsigned*)nullptr, (int)n, st);
  WALLOC(tmp, std::max(std::max(b1, b2), std::max(b3, b4)) + 256);
  size_t tb = std::max(std::max(b1, b2), std::max(b3, b4)) + 256;
  cub::DeviceRadixSort::SortPairs(tmp.p, tb, kz.as<unsigned>(), kz_s.as<unsigned>(), pos.as<unsigned>(), pos_s.as<unsigned>(), (int)n, 0, 32, st);
  weld_gather_kernel<<<g, 256, 0, st>>>(kxy.as<u64>(), pos_s.as<unsigned>(), n, kxy_g.as<u64>());
  cub::DeviceRadixSort::SortPairs(tmp.p, tb, kxy_g.as<u64>(), kxy_s.as<u64>(), pos_s.as<unsigned>(), pos2.as<unsigned>(), (int)n, 0, 64, st);
  // the sorted-key scratch is dead from here on: reuse it
  WALLOC(head, (size_t)n * 4); WALLOC(run, (size_t)n * 4); WALLOC(rep, (size_t)n * 4); WALLOC(isrep, (size_t)n * 4); WALLOC(id, (size_t)n * 4);
  weld_heads_kernel<<<g, 256, 0, st>>>(kxy.as<u64>(), kz.as<unsigned>(), pos2.as<unsigned>(), n, head.as<unsigned>());
  cub::DeviceScan::InclusiveScan(tmp.p, tb, head.as<unsigned>(), run.as<unsigned>(), MaxOp(), (int)n, st);
  weld_reps_kernel<<<g, 256, 0, st>>>(pos2.as<unsigned>(), run.as<unsigned>(), n, rep.as<unsigned>(), isrep.as<unsigned>());
  cub::DeviceScan::ExclusiveSum(tmp.p, tb, isrep.as<unsigned>(), id.as<unsigned>(), (int)n, st);
  unsigned last_id = 0, last_flag = 0;
  CK(cudaMemcpyAsync(&last_id, id.as<unsigned>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last_flag, isrep.as<unsigned>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const unsigned nv = last_id + last_flag;
  WALLOC(dverts, (size_t)nv * sizeof(vh_vertex)); WALLOC(dfaces, (size_t)n * 4);
#undef WALLOC
  weld_emit_kernel<<<g, 256, 0, st>>>(d_soup, n, rep.as<unsigned>(), isrep.as<unsigned>(), id.as<unsigned>(), e->P.vox_size, dverts.as<vh_vertex>(),
                                      dfaces.as<int32_t>());
  verts.resize(nv); faces.resize(n);
  CK(cudaMemcpyAsync(verts.data(), dverts.p, (size_t)nv * sizeof(vh_vertex), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(faces.data(), dfaces.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return VH_OK;
}
