// vh_map_api.cu — the spatial hash as a standalone C-ABI object (vh_map_*), for callers that used
// vhashing::HashTable<int3, ...> directly (/root/reference/include/vhashing.h:627-826). Bulk, device-side
// equivalents of AllocKeys / find / erase (vhashing.h:531-603, :140-142, :278-290); include/vhashing.h of this
// repo wraps the same MapView for use inside callers' own kernels.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/vh_c.h"
#include "../../include/vh_map.cuh"

using namespace vh;

struct vh_map {
  MapView V;
  uint32_t capacity;
  int device;
  bool erased_any;
  u64* d_keys; int* d_out; size_t tmp_cap;
};

// errors of this translation unit are reported through the same thread-local string as the engine's
extern "C" int vh_set_error_(int code, const char* msg);

#define MCK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess) { char b[256]; snprintf(b, sizeof b, "CUDA Error: %s (%s)", cudaGetErrorString(_e), #call); return vh_set_error_(VH_ERR_CUDA, b); } \
  } while (0)

__global__ void map_init_free_list(int* fl, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fl[i] = n - 1 - i;
}

// pass 1 of insert: claim entries and pop pool slots (warp-aggregated)
__global__ void map_insert_kernel(MapView V, const u64* __restrict__ keys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const u64 key = i < n ? keys[i] : KEY_EMPTY;
  const unsigned same = __match_any_sync(0xffffffffu, key);
  const bool leader = key != KEY_EMPTY && lane == __ffs(same) - 1;
  int entry = -1;
  bool claimed = false;
  if (leader) entry = map_claim(V, key, claimed);
  map_assign_slots(V, 0xffffffffu, claimed, entry, key);
}
// pass 2 / find: resolve slots
__global__ void map_find_kernel(MapView V, const u64* __restrict__ keys, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = map_find(V, keys[i]);
  out[i] = e >= 0 ? V.slots[e] : -1;
}
// erase: tombstone the entry and push its slot back on the free list; duplicates in the batch erase once
__global__ void map_erase_kernel(MapView V, const u64* __restrict__ keys, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 key = keys[i];
  int erased = 0;
  const int e = map_find(V, key);
  if (e >= 0 && atomicCAS(&V.keys[e], key, KEY_TOMB) == key) {
    const int slot = V.slots[e];
    V.slots[e] = -1;
    if (slot >= 0) map_release_slot(V, slot);
    erased = 1;
  }
  out[i] = erased;
}
// live keys by table scan (used once anything was erased: key_heap then holds stale keys)
__global__ void map_scan_keys_kernel(MapView V, u64* __restrict__ out, int* __restrict__ n_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > V.mask) return;
  const u64 k = V.keys[i];
  if (k != KEY_EMPTY && k != KEY_TOMB) out[atomicAdd(n_out, 1)] = k;
}

static int ensure_tmp(vh_map* m, size_t n) {
  if (n <= m->tmp_cap) return VH_OK;
  cudaFree(m->d_keys); cudaFree(m->d_out);
  m->d_keys = nullptr; m->d_out = nullptr; m->tmp_cap = 0;
  MCK(cudaMalloc((void**)&m->d_keys, n * sizeof(u64)));
  MCK(cudaMalloc((void**)&m->d_out, n * sizeof(int)));
  m->tmp_cap = n;
  return VH_OK;
}

static int stage_keys(vh_map* m, const int32_t* xyz, int n) {
  int rc = ensure_tmp(m, (size_t)n);
  if (rc != VH_OK) return rc;
  std::vector<u64> packed((size_t)n);
  for (int i = 0; i < n; i++) {
    if (!key_in_range(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])) return vh_set_error_(VH_ERR_INVALID, "block coordinate outside [-2^20, 2^20)");
    packed[i] = pack_key(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
  MCK(cudaMemcpy(m->d_keys, packed.data(), (size_t)n * sizeof(u64), cudaMemcpyHostToDevice));
  return VH_OK;
}

static int check_flags(vh_map* m) {
  int f = 0;
  MCK(cudaMemcpy(&f, m->V.error_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (f & MAP_TABLE_FULL) return vh_set_error_(VH_ERR_TABLE_FULL, "hash table full");
  if (f & MAP_POOL_FULL) return vh_set_error_(VH_ERR_POOL_FULL, "out of block memory");
  return VH_OK;
}

extern "C" {

int vh_map_create(int num_buckets, int entries_per_bucket, int num_blocks, int device, vh_map** out) {
  if (!out || num_buckets <= 0 || entries_per_bucket <= 0 || num_blocks <= 0) return vh_set_error_(VH_ERR_INVALID, "invalid argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return vh_set_error_(VH_ERR_NO_DEVICE, "no CUDA device: no CPU fallback");
  if (device < 0 || device >= ndev) return vh_set_error_(VH_ERR_INVALID, "device out of range");
  MCK(cudaSetDevice(device));
  uint64_t want = (uint64_t)num_buckets * entries_per_bucket, cap = 1024;
  while (cap < want) cap <<= 1;
  if (cap > (1ull << 31)) return vh_set_error_(VH_ERR_INVALID, "hash table too large");
  vh_map* m = new vh_map();
  m->capacity = (uint32_t)cap; m->device = device; m->erased_any = false; m->d_keys = nullptr; m->d_out = nullptr; m->tmp_cap = 0;
  MapView& V = m->V;
  V.mask = m->capacity - 1; V.num_blocks = num_blocks;
  cudaError_t ce = cudaMalloc((void**)&V.keys, cap * sizeof(u64));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&V.slots, cap * sizeof(int));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&V.free_list, (size_t)num_blocks * sizeof(int));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&V.free_top, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&V.key_heap, (size_t)num_blocks * sizeof(u64));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&V.heap_counter, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&V.error_flag, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMemset(V.keys, 0xFF, cap * sizeof(u64));
  if (ce == cudaSuccess) ce = cudaMemset(V.slots, 0xFF, cap * sizeof(int));
  if (ce == cudaSuccess) ce = cudaMemset(V.heap_counter, 0, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMemset(V.error_flag, 0, sizeof(int));
  if (ce == cudaSuccess) ce = cudaMemcpy(V.free_top, &num_blocks, sizeof(int), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) { map_init_free_list<<<(num_blocks + 255) / 256, 256>>>(V.free_list, num_blocks); ce = cudaDeviceSynchronize(); }
  if (ce != cudaSuccess) { vh_map_destroy(m); return vh_set_error_(VH_ERR_CUDA, cudaGetErrorString(ce)); }
  *out = m;
  return VH_OK;
}

int vh_map_destroy(vh_map* m) {
  if (!m) return VH_OK;
  cudaSetDevice(m->device);
  cudaFree(m->V.keys); cudaFree(m->V.slots); cudaFree(m->V.free_list); cudaFree(m->V.free_top); cudaFree(m->V.key_heap);
  cudaFree(m->V.heap_counter); cudaFree(m->V.error_flag); cudaFree(m->d_keys); cudaFree(m->d_out);
  delete m;
  return VH_OK;
}

int vh_map_insert(vh_map* m, const int32_t* keys_xyz, int n, int32_t* out_slots) {
  if (!m || (n > 0 && !keys_xyz)) return vh_set_error_(VH_ERR_INVALID, "null argument");
  if (n <= 0) return VH_OK;
  MCK(cudaSetDevice(m->device));
  int rc = stage_keys(m, keys_xyz, n);
  if (rc != VH_OK) return rc;
  const int grid = (n + 255) / 256;
  map_insert_kernel<<<grid, 256>>>(m->V, m->d_keys, n);
  map_find_kernel<<<grid, 256>>>(m->V, m->d_keys, n, m->d_out);
  MCK(cudaDeviceSynchronize());
  if (out_slots) MCK(cudaMemcpy(out_slots, m->d_out, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return check_flags(m);
}

int vh_map_find(vh_map* m, const int32_t* keys_xyz, int n, int32_t* out_slots) {
  if (!m || (n > 0 && (!keys_xyz || !out_slots))) return vh_set_error_(VH_ERR_INVALID, "null argument");
  if (n <= 0) return VH_OK;
  MCK(cudaSetDevice(m->device));
  int rc = stage_keys(m, keys_xyz, n);
  if (rc != VH_OK) return rc;
  map_find_kernel<<<(n + 255) / 256, 256>>>(m->V, m->d_keys, n, m->d_out);
  MCK(cudaDeviceSynchronize());
  MCK(cudaMemcpy(out_slots, m->d_out, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return VH_OK;
}

int vh_map_erase(vh_map* m, const int32_t* keys_xyz, int n, int32_t* out_erased) {
  if (!m || (n > 0 && !keys_xyz)) return vh_set_error_(VH_ERR_INVALID, "null argument");
  if (n <= 0) return VH_OK;
  MCK(cudaSetDevice(m->device));
  int rc = stage_keys(m, keys_xyz, n);
  if (rc != VH_OK) return rc;
  map_erase_kernel<<<(n + 255) / 256, 256>>>(m->V, m->d_keys, n, m->d_out);
  MCK(cudaDeviceSynchronize());
  m->erased_any = true;
  if (out_erased) MCK(cudaMemcpy(out_erased, m->d_out, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return VH_OK;
}

int vh_map_keys(vh_map* m, int32_t* out_xyz, int cap, int* n_out) {
  if (!m || !n_out) return vh_set_error_(VH_ERR_INVALID, "null argument");
  MCK(cudaSetDevice(m->device));
  std::vector<u64> keys;
  int n = 0;
  if (!m->erased_any) {
    MCK(cudaMemcpy(&n, m->V.heap_counter, sizeof(int), cudaMemcpyDeviceToHost));
    if (n > m->V.num_blocks) n = m->V.num_blocks;
    keys.resize((size_t)n);
    if (n) MCK(cudaMemcpy(keys.data(), m->V.key_heap, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost));
  } else {
    u64* d_tmp = nullptr; int* d_n = nullptr;
    MCK(cudaMalloc((void**)&d_tmp, (size_t)m->V.num_blocks * sizeof(u64)));
    MCK(cudaMalloc((void**)&d_n, sizeof(int)));
    MCK(cudaMemset(d_n, 0, sizeof(int)));
    map_scan_keys_kernel<<<(m->capacity + 255) / 256, 256>>>(m->V, d_tmp, d_n);
    cudaError_t ce = cudaMemcpy(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost);
    keys.resize((size_t)n);
    if (ce == cudaSuccess && n) ce = cudaMemcpy(keys.data(), d_tmp, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost);
    cudaFree(d_tmp); cudaFree(d_n);
    MCK(ce);
  }
  *n_out = n;
  if (out_xyz)
    for (int i = 0; i < n && i < cap; i++) unpack_key(keys[i], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
  return VH_OK;
}

int vh_map_size(vh_map* m, int* n) {
  if (!m || !n) return vh_set_error_(VH_ERR_INVALID, "null argument");
  MCK(cudaSetDevice(m->device));
  int top = 0;
  MCK(cudaMemcpy(&top, m->V.free_top, sizeof(int), cudaMemcpyDeviceToHost));
  *n = m->V.num_blocks - (top > 0 ? top : 0);
  return VH_OK;
}

int vh_map_get_view(vh_map* m, vh_map_view* out) {
  if (!m || !out) return vh_set_error_(VH_ERR_INVALID, "null argument");
  out->keys = m->V.keys; out->slots = m->V.slots; out->capacity_mask = m->V.mask; out->free_list = m->V.free_list;
  out->free_top = m->V.free_top; out->key_heap = m->V.key_heap; out->heap_counter = m->V.heap_counter;
  out->error_flag = m->V.error_flag; out->num_blocks = m->V.num_blocks;
  return VH_OK;
}

}  // extern "C"
