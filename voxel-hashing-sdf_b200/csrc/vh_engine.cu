// vh_engine.cu — host side of the engine and the C ABI of include/vh_c.h.
//
// Orchestration replaces GpuTsdfGenerator::processFrame (/root/reference/src/tsdf.cu:1485-1598), which per frame
// builds and destroys two hash tables, performs ~20 blocking copies and 17 cudaMalloc/cudaFree pairs. Here a frame is:
// two async H2D copies on an upload stream (double-buffered, overlapping the previous frame's kernels), one 64-byte
// counter reset, and three kernel launches (allocate, integrate, marching cubes) on the compute stream. Nothing is
// allocated, freed or synchronised per frame unless the caller asks for the synchronous entry point.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <vector>

#include "vh_engine_host.h"
#include "vh_params_host.h"

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
extern "C" int vh_set_error_(int code, const char* msg) { g_err = msg ? msg : ""; return code; }

constexpr int VH_MC_COLOR_TILE_DEFAULT = 0;      // set from the measurement in profiles/r02final (both settings run in the GPU test suite)

// ---- frame constants on the host: derive_frame_params (vh_params_host.h)
void setup_frame(vh_engine* e, const float* c2w) {
  derive_frame_params(e->P, e->S, c2w, e->F);
  e->F.frame = (uint32_t)(++e->frames);
}

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* vh_last_error(void) { return g_err.c_str(); }
const char* vh_version(void) { return "vhsdf-b200 0.1 (sm_100a)"; }

int vh_default_params(vh_params* p) {
  if (!p) return fail(VH_ERR_INVALID, "null params");
  memset(p, 0, sizeof(*p));
  p->width = 640; p->height = 480;
  p->fx = p->fy = 577.0f; p->cx = 320.0f; p->cy = 240.0f;        // scene0220_02.yaml:11-14
  p->min_depth = 0.1f; p->max_depth = 10.0f;
  p->vox_size = 0.01f; p->trunc_margin = 0.05f;
  p->voxels_per_block = 8; p->blocks_per_chunk = 8;
  p->dda_stride = 10; p->max_ray_steps = 100;
  p->chunk_radius = 4.0f; p->max_chunk_num = 128;
  p->num_buckets = 1 << 20; p->entries_per_bucket = 4;
  p->pool_blocks = 1 << 20;
  p->use_color = 1; p->mc_per_frame = 1;
  p->device = 0; p->shard_rank = 0; p->shard_count = 1; p->shard_group = 0;
  p->depth_tile_smem = 1;
  p->tri_arena_bytes = 0;
  return VH_OK;
}

static int free_engine(vh_engine* e) {
  if (!e) return VH_OK;
  cudaSetDevice(e->P.device);
  shard_release(e);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->upload) cudaStreamSynchronize(e->upload);
  DeviceView& D = e->D;
  cudaFree(D.map.keys); cudaFree(D.map.slots); cudaFree(D.map.free_list); cudaFree(D.map.free_top); cudaFree(D.map.key_heap);
  cudaFree(e->d_status); cudaFree(D.stamps); cudaFree(D.sdf); cudaFree(D.wgt); cudaFree(D.rgb); cudaFree(D.tile_max); cudaFree(D.sched); cudaFree(D.work); cudaFree(D.neg_count); cudaFree(D.mc_queue); cudaFree(D.mc_ctl); cudaFree(D.visible); cudaFree(D.inbox); cudaFree(D.inbox_count);
  cudaFree(D.arena); cudaFree(e->arena_spare); cudaFree(e->d_scan_in); cudaFree(e->d_scan_out); cudaFree(e->d_scan_tmp);
  cudaFree(D.tri_offset); cudaFree(D.tri_count);
  cudaFree(e->d_full_list); cudaFree(e->d_full_count); cudaFree(e->d_full_off); cudaFree(e->d_full_cnt); cudaFree(e->d_keys_tmp);
  for (int i = 0; i < 2; i++) {
    cudaFree(e->d_depth[i]); cudaFree(e->d_depth16[i]); cudaFree(e->d_rgb[i]); cudaFree(e->d_px[i]);
    if (e->ev_uploaded[i]) cudaEventDestroy(e->ev_uploaded[i]);
    if (e->ev_rgb[i]) cudaEventDestroy(e->ev_rgb[i]);
    if (e->ev_consumed[i]) cudaEventDestroy(e->ev_consumed[i]);
  }
  for (int i = 0; i < 6; i++) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  if (e->ev_mapped_read) cudaEventDestroy(e->ev_mapped_read);
  if (e->h_block) cudaFreeHost(e->h_block);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->upload) cudaStreamDestroy(e->upload);
  delete e;
  return VH_OK;
}

__global__ void init_free_list_kernel(int* free_list, int n) {
  // slot s is handed out in ascending order: the stack top holds slot 0
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) free_list[i] = n - 1 - i;
}

static int reset_map(vh_engine* e) {
  DeviceView& D = e->D;
  const int nb = e->P.pool_blocks;
  CK(cudaMemsetAsync(D.map.keys, 0xFF, (size_t)e->capacity * sizeof(u64), e->stream));
  CK(cudaMemsetAsync(D.map.slots, 0xFF, (size_t)e->capacity * sizeof(int), e->stream));
  CK(cudaMemsetAsync(D.stamps, 0, (size_t)e->capacity * sizeof(uint32_t), e->stream));
  CK(cudaMemsetAsync(D.sdf, 0, (size_t)nb * BLOCK_VOX * sizeof(float), e->stream));      // Voxel(): sdf = 0, weight = 0 (tsdf.cuh:126-128)
  CK(cudaMemsetAsync(D.wgt, 0, (size_t)nb * BLOCK_VOX * sizeof(float), e->stream));
  if (D.rgb) CK(cudaMemsetAsync(D.rgb, 0, (size_t)nb * BLOCK_VOX * sizeof(uchar4), e->stream));
  CK(cudaMemsetAsync(D.neg_count, 0, (size_t)nb * sizeof(int), e->stream));
  CK(cudaMemsetAsync(D.mc_ctl, 0, 2 * sizeof(McQueueCtl), e->stream));
  CK(cudaMemsetAsync(D.inbox_count, 0, 4 * sizeof(int), e->stream));
  CK(cudaMemsetAsync(D.tri_offset, 0, (size_t)nb * sizeof(unsigned long long), e->stream));
  CK(cudaMemsetAsync(D.tri_count, 0, (size_t)nb * sizeof(int), e->stream));
  CK(cudaMemsetAsync(e->d_status, 0, sizeof(DeviceStatus), e->stream));
  init_free_list_kernel<<<(nb + 255) / 256, 256, 0, e->stream>>>(D.map.free_list, nb);
  CK(cudaMemcpyAsync(D.map.free_top, &nb, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->tombstones = 0;
  e->frames = 0; e->updates_total = 0; e->max_tris_per_frame = 0; e->known_arena_top = 0; e->frames_in_flight = 0; e->compactions = 0; e->forced_syncs = 0; e->integrate_launches = 0; e->weight_bound_bias = e->weight_bound_env;
  memset(e->h_block, 0, sizeof(*e->h_block));
  return VH_OK;
}

int vh_create(const vh_params* p, vh_engine** out) {
  if (!p || !out) return fail(VH_ERR_INVALID, "null argument");
  *out = nullptr;
  if (p->voxels_per_block != VPB) return fail(VH_ERR_INVALID, "voxels_per_block must be %d (got %d)", VPB, p->voxels_per_block);
  if (p->width <= 0 || p->height <= 0 || p->vox_size <= 0 || p->trunc_margin <= 0 || p->dda_stride <= 0 || p->max_ray_steps <= 0 ||
      p->blocks_per_chunk <= 0 || p->num_buckets <= 0 || p->entries_per_bucket <= 0 || p->pool_blocks <= 0 || p->shard_count <= 0 ||
      p->shard_rank < 0 || p->shard_rank >= p->shard_count)
    return fail(VH_ERR_INVALID, "invalid parameter value");
  if (!(p->trunc_margin > 1e-15f && p->trunc_margin < 1e15f && p->vox_size > 1e-15f && p->vox_size < 1e15f))
    return fail(VH_ERR_INVALID, "trunc_margin / vox_size outside the range the integrate kernel's divisions are validated for");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(VH_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (p->device < 0 || p->device >= ndev) return fail(VH_ERR_INVALID, "device %d out of range (%d devices)", p->device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, p->device));
  if (prop.major != 10) return fail(VH_ERR_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a (B200) only", p->device, prop.major, prop.minor);
  CK(cudaSetDevice(p->device));

  vh_engine* e = new vh_engine;
  e->P = *p;
  e->num_sms = prop.multiProcessorCount;
  StaticParams& S = e->S;
  derive_static_params(*p, S);
  { const char* v = getenv("VH_INTEGRATE_VERIFY"); S.verify = (v && v[0] == '1') ? 1 : 0; }
  { const char* v = getenv("VH_INTEGRATE_CTAS"); S.integrate_ctas_per_sm = (v && v[0] >= '3' && v[0] <= '6') ? v[0] - '0' : 0; }   // tuning knob; 0 = the kernel's default (direct 3 [or 4], staged 4 [or 5, 6])
  { const char* v = getenv("VH_INTEGRATE_EXACT_COLOR"); e->weight_bound_env = e->weight_bound_bias = (v && v[0] == '1') ? 1u << 20 : 0u; }
  { const char* v = getenv("VH_INTEGRATE_CULL"); S.integrate_cull = (v && v[0] == '0') ? 0 : 1; }
  { const char* v = getenv("VH_INTEGRATE_PARTS"); e->integrate_parts_forced = (v && (v[0] == '1' || v[0] == '2')) ? v[0] - '0' : 0; S.integrate_parts = 2; }
  { const char* v = getenv("VH_INTEGRATE_REV"); S.integrate_rev = (v && v[0] == '1') ? 1 : 2; }      // 2 (default) = integrate_kernel_staged, 1 = integrate_kernel_direct
  { // allocation: 2 = ray_keys_kernel + insert_keys_kernel (sharded maps: rays split across the GPUs; long ray step caps), 0 = the one-kernel
    // form (a few us faster at the reference's 100-step cap on one GPU); VH_ALLOC_REV overrides
    const char* v = getenv("VH_ALLOC_REV");
    S.alloc_rev = (v && (v[0] == '0' || v[0] == '2')) ? v[0] - '0' : ((p->shard_count > 1 || p->max_ray_steps > 256) ? 2 : 0);
  }
  { const char* v = getenv("VH_MC_COLOR_TILE"); S.mc_rev = v ? (v[0] == '1' ? 1 : 0) : VH_MC_COLOR_TILE_DEFAULT; }   // mesh kernel: colour tile through shared memory
  static_assert(sizeof(StaticParams) % 16 == 0, "keep the FrameParams behind StaticParams 16-byte aligned in the kernels' parameter blocks");

  uint64_t want = (uint64_t)p->num_buckets * (uint64_t)p->entries_per_bucket;
  uint64_t cap = 1024;
  while (cap < want) cap <<= 1;
  if (cap > (1ull << 31)) { delete e; return fail(VH_ERR_INVALID, "hash table too large"); }
  e->capacity = (uint32_t)cap;

  DeviceView& D = e->D;
  memset(&D, 0, sizeof(D));
  const size_t nb = (size_t)p->pool_blocks;
  const size_t rays = (size_t)S.nrx * S.nry;
  D.list_cap = (int)std::max(rays * (size_t)S.max_steps, nb);
  uint64_t arena_bytes = p->tri_arena_bytes ? p->tri_arena_bytes : (1ull << 30);
  D.arena_cap = arena_bytes / sizeof(vh_triangle);
#define ALLOC(ptr, bytes)                                                                                           \
  do {                                                                                                              \
    cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                                                           \
    if (_e != cudaSuccess) { cudaGetLastError(); /* not sticky: keep it from surfacing in a later, unrelated call */ \
      int rc = fail(VH_ERR_CUDA, "CUDA Error: cudaMalloc(%zu bytes) for %s: %s", (size_t)(bytes), #ptr, cudaGetErrorString(_e)); free_engine(e); return rc; } \
  } while (0)
  ALLOC(D.map.keys, cap * sizeof(u64));
  ALLOC(D.map.slots, cap * sizeof(int));
  ALLOC(D.stamps, cap * sizeof(uint32_t));
  ALLOC(D.map.free_list, nb * sizeof(int));
  ALLOC(D.map.free_top, sizeof(int));
  ALLOC(D.map.key_heap, nb * sizeof(u64));
  ALLOC(e->d_status, sizeof(DeviceStatus));
  ALLOC(D.sdf, nb * BLOCK_VOX * sizeof(float));
  ALLOC(D.wgt, nb * BLOCK_VOX * sizeof(float));
  if (S.use_color) ALLOC(D.rgb, nb * BLOCK_VOX * sizeof(uchar4));
  ALLOC(D.neg_count, nb * sizeof(int));
  ALLOC(D.sched, 9 * 32 * sizeof(int));
  ALLOC(D.work, (size_t)D.list_cap * sizeof(uint4));
  ALLOC(D.tile_max, (size_t)((p->width + 15) / 16) * ((p->height + 15) / 16) * sizeof(float));
  ALLOC(D.mc_queue, (size_t)D.list_cap * sizeof(McWork));
  ALLOC(D.mc_ctl, 2 * sizeof(McQueueCtl));
  ALLOC(D.visible, (size_t)D.list_cap * sizeof(int));
  D.inbox_cap = (int)std::max<size_t>(rays * (size_t)S.max_steps, 1024);
  ALLOC(D.inbox, 2 * (size_t)D.inbox_cap * sizeof(u64));
  ALLOC(D.inbox_count, 4 * sizeof(int));
  D.inbox_done = D.inbox_count + 2;
  ALLOC(D.arena, D.arena_cap * sizeof(vh_triangle));
  if (p->mc_per_frame) ALLOC(e->arena_spare, D.arena_cap * sizeof(vh_triangle));      // compaction target; the two swap roles
  ALLOC(D.tri_offset, nb * sizeof(unsigned long long));
  ALLOC(D.tri_count, nb * sizeof(int));
  const size_t npx = (size_t)p->width * p->height;
  for (int i = 0; i < 2; i++) {
    ALLOC(e->d_depth[i], npx * sizeof(float));
    ALLOC(e->d_rgb[i], npx * 3);
    ALLOC(e->d_px[i], (npx + 1) * sizeof(uint2));    // + the sentinel record {depth 0, rgb 0} integrate_kernel_r1 reads for pixels outside the image
    if (cudaMemset(e->d_px[i] + npx, 0, sizeof(uint2)) != cudaSuccess) { int rc = fail(VH_ERR_CUDA, "CUDA Error: cudaMemset of the frame sentinel"); free_engine(e); return rc; }
  }
#undef ALLOC
  D.map.mask = e->capacity - 1;
  D.map.num_blocks = p->pool_blocks;
  D.mc_parity = &e->mc_parity;
  D.counters = &e->d_status->c;
  D.map.error_flag = &e->d_status->map_error;
  D.engine_error = &e->d_status->engine_error;
  D.map.heap_counter = &e->d_status->heap_counter;
  D.overflow_frame = &e->d_status->overflow_frame;
  D.arena_top = &e->d_status->arena_top;
  D.updates_total = &e->d_status->updates_total;
  bool ok = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&e->upload, cudaStreamNonBlocking) == cudaSuccess &&
            cudaHostAlloc((void**)&e->h_block, sizeof(*e->h_block), cudaHostAllocDefault) == cudaSuccess;
  for (int i = 0; i < 2 && ok; i++)
    ok = cudaEventCreateWithFlags(&e->ev_uploaded[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_rgb[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 6 && ok; i++) ok = cudaEventCreate(&e->ev[i]) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&e->ev_mapped_read, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { free_engine(e); return fail(VH_ERR_CUDA, "CUDA Error: stream/event creation failed"); }
  {   // the status block is published by a kernel into mapped pinned memory (see publish_status_kernel); VH_STATUS_PUBLISH=0 = D2H copy node instead
    const char* v = getenv("VH_STATUS_PUBLISH");
    void* dp = nullptr;
    if (!(v && v[0] == '0')) { if (cudaHostGetDevicePointer(&dp, e->h_block, 0) == cudaSuccess) e->h_block_dev = static_cast<DeviceStatus*>(dp); else cudaGetLastError(); }
  }
  upload_mc_tables();
  int rc = reset_map(e);
  if (rc != VH_OK) { free_engine(e); return rc; }
  e->cur_depth = e->d_depth[0]; e->cur_rgb = e->d_rgb[0];
  *out = e;
  return VH_OK;
}

int vh_destroy(vh_engine* e) { return free_engine(e); }

int vh_reset(vh_engine* e) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  if (e->shard) { int rc = shard_barrier(e); if (rc != VH_OK) return rc; }     // sharded map (collective): no peer is still meshing against these voxels
  CK(cudaStreamSynchronize(e->stream));
  int rc = reset_map(e);
  if (rc != VH_OK || !e->shard) return rc;
  // ... and nobody starts the next frame before EVERY GPU has finished resetting: a fast peer's ray_keys_kernel would otherwise
  // add keys to this GPU's inbox (and bump its counter) ahead of the memsets above
  rc = shard_barrier(e);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return VH_OK;
}

// ---- triangle arena maintenance -----------------------------------------------------------------
__global__ void compact_copy_kernel(const vh_triangle* __restrict__ src, vh_triangle* __restrict__ dst, const unsigned long long* __restrict__ old_off,
                                    const unsigned long long* __restrict__ new_off, const int* __restrict__ cnt, int n) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n || cnt[w] <= 0) return;
  const uint4* s = reinterpret_cast<const uint4*>(src + old_off[w]);
  uint4* d = reinterpret_cast<uint4*>(dst + new_off[w]);
  for (int i = lane; i < cnt[w] * 3; i += 32) d[i] = s[i];
}
__global__ void widen_counts_kernel(const int* cnt, unsigned long long* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cnt[i] > 0 ? (unsigned long long)cnt[i] : 0ull;
}

// Drop superseded triangles (blocks re-meshed in later frames) and make room for `need` more. Stream must be idle.
// The live ranges are copied into the spare arena (same size, allocated with the engine) and the two swap roles, so a
// steady-state compaction allocates nothing; only growth (live set above half an arena) reallocates.
static int compact_arena(vh_engine* e, unsigned long long need) {
  DeviceView& D = e->D;
  const int nb = e->P.pool_blocks;
  if (!e->d_scan_in) {
    CK(cudaMalloc((void**)&e->d_scan_in, (size_t)nb * sizeof(unsigned long long)));
    CK(cudaMalloc((void**)&e->d_scan_out, ((size_t)nb + 1) * sizeof(unsigned long long)));
    cub::DeviceScan::ExclusiveSum(nullptr, e->scan_tmp_bytes, e->d_scan_in, e->d_scan_out, nb, e->stream);
    CK(cudaMalloc(&e->d_scan_tmp, e->scan_tmp_bytes));
  }
  unsigned long long *d_wide = e->d_scan_in, *d_new = e->d_scan_out;
  widen_counts_kernel<<<(nb + 255) / 256, 256, 0, e->stream>>>(D.tri_count, d_wide, nb);
  cub::DeviceScan::ExclusiveSum(e->d_scan_tmp, e->scan_tmp_bytes, d_wide, d_new, nb, e->stream);
  unsigned long long last_off = 0, last_cnt = 0;
  CK(cudaMemcpyAsync(&last_off, d_new + nb - 1, sizeof(last_off), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(&last_cnt, d_wide + nb - 1, sizeof(last_cnt), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const unsigned long long live = last_off + last_cnt;
  unsigned long long new_cap = D.arena_cap;
  while (live + need > new_cap / 2) new_cap *= 2;      // keep at least half the arena free after compaction
  vh_triangle* fresh = nullptr;
  if (new_cap == D.arena_cap && e->arena_spare) {
    fresh = e->arena_spare;
    e->arena_spare = nullptr;
  } else {
    cudaFree(e->arena_spare); e->arena_spare = nullptr;
    cudaError_t ce = cudaMalloc((void**)&fresh, new_cap * sizeof(vh_triangle));
    if (ce != cudaSuccess) return fail(VH_ERR_ARENA_FULL, "triangle arena cannot grow to %llu triangles: %s", new_cap, cudaGetErrorString(ce));
  }
  compact_copy_kernel<<<(nb + 7) / 8, 256, 0, e->stream>>>(D.arena, fresh, D.tri_offset, d_new, D.tri_count, nb);
  CK(cudaMemcpyAsync(D.tri_offset, d_new, (size_t)nb * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(D.arena_top, &live, sizeof(live), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (new_cap == D.arena_cap) e->arena_spare = D.arena;     // the old arena becomes the spare
  else {
    cudaFree(D.arena);
    if (cudaMalloc((void**)&e->arena_spare, new_cap * sizeof(vh_triangle)) != cudaSuccess) { e->arena_spare = nullptr; cudaGetLastError(); }   // next compaction allocates
  }
  D.arena = fresh; D.arena_cap = new_cap;
  e->known_arena_top = live;
  e->compactions++;
  return VH_OK;
}

// Work items of the staged integrate kernel for the coming launch: whole blocks when the last completed frame kept many blocks
// after the whole-block discard (room-scale frames: ~470 k), x-halves otherwise (more items for the warps to share, smaller
// staging buffers, 5 CTAs per SM). A performance choice only — results do not depend on it. The pinned status block is
// refreshed by every frame, so it can be read without a sync.
void pick_integrate_parts(vh_engine* e) {
  if (e->integrate_parts_forced) { e->S.integrate_parts = e->integrate_parts_forced; return; }
  const volatile DeviceStatus* hb = e->h_block;
  const long long kept = (long long)hb->c.visible_count - (long long)hb->c.pad[1];
  e->S.integrate_parts = kept >= 160000 ? 1 : 2;
}

// ---- frame pipeline -----------------------------------------------------------------------------
int enqueue_stages(vh_engine* e, bool do_alloc, cudaEvent_t depth_ready, cudaEvent_t rgb_ready, const float* host_depth_mapped) {
  DeviceView& D = e->D;
  uint2* px = e->d_px[e->px_ring & 1]; e->px_ring++;
  if (do_alloc && (rgb_ready || host_depth_mapped)) {
    // Host frames: the ray pass needs only ~3,000 depth samples. With a pinned (device-mapped) caller buffer it reads them
    // straight from host memory and runs while BOTH images are still uploading; otherwise it waits for the depth upload
    // and overlaps the colour upload only.
    if (!host_depth_mapped && depth_ready) CK(cudaStreamWaitEvent(e->stream, depth_ready, 0));
    CK(cudaMemsetAsync(D.counters, 0, sizeof(FrameCounters), e->stream));
    launch_alloc_visible(e->S, e->F, host_depth_mapped ? host_depth_mapped : e->cur_depth, D, e->stream);
    if (host_depth_mapped) {      // the caller's buffer is being read in place on THIS stream: vh_wait_uploads must cover it too
      CK(cudaEventRecord(e->ev_mapped_read, e->stream));
      e->mapped_read_pending = true;
    }
    if (host_depth_mapped && depth_ready) CK(cudaStreamWaitEvent(e->stream, depth_ready, 0));
    if (rgb_ready) CK(cudaStreamWaitEvent(e->stream, rgb_ready, 0));
    launch_pack_frame(e->S, e->cur_depth, e->cur_rgb, px, D.tile_max, D.sched, D.counters, e->F.frame, e->stream, 1);
  } else {
    if (depth_ready) CK(cudaStreamWaitEvent(e->stream, depth_ready, 0));
    // first kernel of the frame: packs {depth, rgb} records for integrate and resets the frame's counters
    launch_pack_frame(e->S, e->cur_depth, e->cur_rgb, px, D.tile_max, D.sched, do_alloc ? D.counters : nullptr, e->F.frame, e->stream);
    if (do_alloc) launch_alloc_visible(e->S, e->F, e->cur_depth, D, e->stream);
  }
  CK(cudaEventRecord(e->ev[5], e->stream));
  launch_cull_list(e->S, e->F, D, e->num_sms, e->stream);          // integrate's work list: visible blocks minus the whole-block discards
  CK(cudaEventRecord(e->ev[2], e->stream));
  e->S.weight_bound = ++e->integrate_launches + e->weight_bound_bias;
  pick_integrate_parts(e);
  launch_integrate(e->S, e->F, px, e->cur_rgb != nullptr, D, e->num_sms, e->stream);
  CK(cudaEventRecord(e->ev[3], e->stream));
  if (e->P.mc_per_frame)
    launch_marching_cubes(e->S, e->F, D, D.visible, &D.counters->visible_count, 0, D.tri_offset, D.tri_count, e->num_sms, e->stream);
  CK(cudaEventRecord(e->ev[4], e->stream));
  return VH_OK;
}

// The status block at the end of a frame (VH_STATUS_PUBLISH=0 falls back to a D2H copy node): one warp stores
// the 128-byte block into the mapped pinned host block instead of a D2H copy node, so that the frame ends without a
// kernel -> copy engine -> kernel hand-off and never queues behind the next frame's upload. Word 0 holds the frame stamp
// the host polls without a sync (make_room); it is stored last, after a system-scope fence.
__global__ void publish_status_kernel(const DeviceStatus* __restrict__ src, DeviceStatus* dst_mapped) {
  static_assert(sizeof(DeviceStatus) == 128, "one 8-byte word per lane of half a warp");
  const int lane = threadIdx.x;
  const unsigned long long* s = reinterpret_cast<const unsigned long long*>(src);
  volatile unsigned long long* d = reinterpret_cast<volatile unsigned long long*>(dst_mapped);
  unsigned long long v = 0;
  if (lane < 16) v = __ldcg(s + lane);
  if (lane > 0 && lane < 16) d[lane] = v;
  __threadfence_system();
  __syncwarp();
  if (lane == 0) { d[0] = v; __threadfence_system(); }
}

int enqueue_readback(vh_engine* e) {
  if (e->h_block_dev) {
    publish_status_kernel<<<1, 32, 0, e->stream>>>(e->d_status, e->h_block_dev);
    CK(cudaGetLastError());
    return VH_OK;
  }
  CK(cudaMemcpyAsync(e->h_block, e->d_status, sizeof(DeviceStatus), cudaMemcpyDeviceToHost, e->stream));
  return VH_OK;
}

// after a sync: fold the read-back block into host state, surface device-side errors, repair arena overflow
int finish_sync(vh_engine* e) {
  auto* hb = e->h_block;
  e->frames_in_flight = 0;
  e->known_arena_top = hb->arena_top;
  e->max_tris_per_frame = std::max<uint64_t>(e->max_tris_per_frame, hb->c.triangles);
  if (hb->map_error & MAP_TABLE_FULL) return fail(VH_ERR_TABLE_FULL, "hash table full (%u entries): raise num_buckets/entries_per_bucket", e->capacity);
  if (hb->map_error & MAP_POOL_FULL) return fail(VH_ERR_POOL_FULL, "out of block memory: pool of %d voxel blocks exhausted, raise pool_blocks", e->P.pool_blocks);
  if (hb->engine_error & 8) return fail(VH_ERR_CUDA, "CUDA Error: a peer GPU of the sharded map did not reach a frame barrier within 4 s");
  if (hb->engine_error & 2) return fail(VH_ERR_CUDA, "CUDA Error: integrate met a value outside the validated range of its division sequence");
  if (hb->engine_error & 1) {
    // marching cubes ran out of arena. Compact/grow; if only the last frame was hit, redo it (it only reads voxels + stamps).
    const uint32_t first_bad = hb->overflow_frame, last = e->F.frame;
    CK(cudaMemsetAsync(e->D.engine_error, 0, 2 * sizeof(int), e->stream));      // engine_error + overflow_frame are adjacent
    CK(cudaStreamSynchronize(e->stream));
    int rc = compact_arena(e, std::max<unsigned long long>(hb->c.triangles * 2, 1ull << 20));
    if (rc != VH_OK) return rc;
    if (e->shard)
      return fail(VH_ERR_ARENA_FULL, "triangle arena overflowed in frame %u of a sharded map (a repair would need every GPU): raise tri_arena_bytes "
                  "(arena grown to %llu triangles)", first_bad, (unsigned long long)e->D.arena_cap);
    if (first_bad != last)
      return fail(VH_ERR_ARENA_FULL, "triangle arena overflowed in frame %u while frames up to %u were in flight: blocks meshed in between are "
                  "incomplete until seen again; raise tri_arena_bytes or call vh_sync more often (arena grown to %llu triangles)",
                  first_bad, last, (unsigned long long)e->D.arena_cap);
    CK(cudaMemsetAsync(&e->D.counters->triangles, 0, sizeof(unsigned long long), e->stream));
    launch_marching_cubes(e->S, e->F, e->D, e->D.visible, &e->D.counters->visible_count, 0, e->D.tri_offset, e->D.tri_count, e->num_sms, e->stream);
    rc = enqueue_readback(e);
    if (rc != VH_OK) return rc;
    CK(cudaStreamSynchronize(e->stream));
    if (hb->engine_error & 1) return fail(VH_ERR_ARENA_FULL, "triangle arena overflow persists after growth");
    e->known_arena_top = hb->arena_top;
  }
  return VH_OK;
}

// Keep enough arena head-room for the frames that are in flight. The pinned status block is refreshed by every
// frame's read-back copy, so it can be read without a sync: it tells which frame has completed, where the arena top
// was then and how many triangles a frame produces.
int make_room(vh_engine* e) {
  if (!e->P.mc_per_frame) return VH_OK;
  const volatile DeviceStatus* hb = e->h_block;
  const uint64_t seen_frame = hb->c.frame, seen_top = hb->arena_top, seen_tris = hb->c.triangles;
  e->max_tris_per_frame = std::max<uint64_t>(e->max_tris_per_frame, seen_tris);
  const uint64_t top = std::max<uint64_t>(e->known_arena_top, seen_top);
  const uint64_t behind = e->frames > seen_frame ? e->frames - seen_frame : 0;           // enqueued, not known to be complete
  const unsigned long long per_frame = std::max<unsigned long long>(e->max_tris_per_frame * 2, 1ull << 18);
  const unsigned long long projected = top + per_frame * (unsigned long long)(behind + 2);
  if (projected <= e->D.arena_cap && behind < 256) return VH_OK;
  CK(cudaStreamSynchronize(e->stream));
  e->forced_syncs++;
  int rc = finish_sync(e);
  if (rc != VH_OK) return rc;
  if (e->known_arena_top + per_frame * 2 <= e->D.arena_cap) return VH_OK;
  return compact_arena(e, per_frame * 2);
}

// u16 depth (PNG millimetres, SaveFrame.cpp:174-180) -> f32 metres on the device: convertTo(CV_32FC1) then *= scale,
// i.e. the float value of the sample times the double scale, rounded once to float
__global__ void depth_u16_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int n, double scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __double2float_rn(__dmul_rn((double)(float)in[i], scale));
}

static int integrate_common(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w, bool host_inputs,
                            const uint16_t* depth_u16 = nullptr, double depth_scale = 0.0) {
  if (!e || (!depth && !depth_u16) || !c2w) return fail(VH_ERR_INVALID, "null argument");
  CK(cudaSetDevice(e->P.device));
  int rc = make_room(e);
  if (rc != VH_OK) return rc;
  const size_t npx = (size_t)e->P.width * e->P.height;
  CK(cudaEventRecord(e->ev[0], e->stream));
  if (host_inputs) {
    const int b = e->ring & 1; e->ring++;
    if (e->buf_used[b]) CK(cudaStreamWaitEvent(e->upload, e->ev_consumed[b], 0));     // previous reader of this buffer is done
    if (depth_u16) {          // half the bytes over PCIe; converted on the upload stream, so it overlaps the previous frame too
      if (!e->d_depth16[b]) CK(cudaMalloc((void**)&e->d_depth16[b], npx * sizeof(uint16_t)));
      CK(cudaMemcpyAsync(e->d_depth16[b], depth_u16, npx * sizeof(uint16_t), cudaMemcpyHostToDevice, e->upload));
      depth_u16_to_f32_kernel<<<(int)((npx + 255) / 256), 256, 0, e->upload>>>(e->d_depth16[b], e->d_depth[b], (int)npx, depth_scale);
    } else {
      CK(cudaMemcpyAsync(e->d_depth[b], depth, npx * sizeof(float), cudaMemcpyHostToDevice, e->upload));
    }
    CK(cudaEventRecord(e->ev_uploaded[b], e->upload));
    const bool with_rgb = rgb && e->S.use_color;
    if (with_rgb) {
      CK(cudaMemcpyAsync(e->d_rgb[b], rgb, npx * 3, cudaMemcpyHostToDevice, e->upload));
      CK(cudaEventRecord(e->ev_rgb[b], e->upload));
    }
    // A pinned caller buffer is visible to the GPU at its mapped address (UVA): the ray pass can read its ~3,000 depth
    // samples from there and start before the upload has finished. That pays only when the upload is on the critical
    // path, i.e. nothing is in flight (the synchronous call). With frames in flight the upload already overlaps the
    // previous frame's kernels, and mapped reads would queue behind the next frame's DMA on the PCIe link (measured with
    // vh_integrate_async on config 2: 3,424 frames/s with mapped reads against 3,686 for the u16 route, which has none).
    // "Nothing in flight" is read from the pinned status block: the stamp of the last completed frame.
    const float* mapped = nullptr;
    const bool gpu_idle = e->frames_in_flight == 0 || static_cast<const volatile DeviceStatus*>(e->h_block)->c.frame >= e->frames;
    if (gpu_idle) {
      cudaPointerAttributes pa;
      if (depth && cudaPointerGetAttributes(&pa, depth) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) mapped = static_cast<const float*>(pa.devicePointer);
      cudaGetLastError();
    }
    e->cur_depth = e->d_depth[b];
    e->cur_rgb = (rgb && e->S.use_color) ? e->d_rgb[b] : nullptr;
    e->buf_used[b] = true;
    CK(cudaEventRecord(e->ev[1], e->stream));
    setup_frame(e, c2w);
    const int keep = e->S.use_color;   // colour only when the caller supplied an image
    e->S.use_color = e->cur_rgb ? keep : 0;
    rc = enqueue_stages(e, true, e->ev_uploaded[b], with_rgb ? e->ev_rgb[b] : nullptr, mapped);
    e->S.use_color = keep;
    if (rc != VH_OK) return rc;
    CK(cudaEventRecord(e->ev_consumed[b], e->stream));
  } else {
    e->cur_depth = depth;
    e->cur_rgb = e->S.use_color ? rgb : nullptr;
    CK(cudaEventRecord(e->ev[1], e->stream));
    setup_frame(e, c2w);
    const int keep = e->S.use_color;
    e->S.use_color = e->cur_rgb ? keep : 0;
    rc = enqueue_stages(e, true, nullptr, nullptr, nullptr);
    e->S.use_color = keep;
    if (rc != VH_OK) return rc;
  }
  e->frames_in_flight++;
  return enqueue_readback(e);
}

int vh_integrate_async(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  return integrate_common(e, depth, rgb, c2w, true);
}

int vh_integrate_device(vh_engine* e, const float* d_depth, const uint8_t* d_rgb, const float* c2w) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  return integrate_common(e, d_depth, d_rgb, c2w, false);
}

// depth as the u16 samples of the reference's depth PNGs (millimetres when depth_scale = 0.001, the factor frameLoad
// applies, SaveFrame.cpp:180): uploaded as 2 bytes per pixel and converted on the GPU. Asynchronous like vh_integrate_async.
int vh_integrate_u16_async(vh_engine* e, const uint16_t* depth_u16, double depth_scale, const uint8_t* rgb, const float* c2w) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  return integrate_common(e, nullptr, rgb, c2w, true, depth_u16, depth_scale);
}

int vh_wait_uploads(vh_engine* e) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->upload));
  // a ray pass that read the caller's pinned depth buffer in place (synchronous route from an idle GPU) runs on the compute
  // stream, not on the upload stream: the buffers are reusable only once it has finished as well
  if (e->mapped_read_pending) { CK(cudaEventSynchronize(e->ev_mapped_read)); e->mapped_read_pending = false; }
  return VH_OK;
}

int vh_sync(vh_engine* e) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

int vh_integrate(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  int rc = integrate_common(e, depth, rgb, c2w, true);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

// ---- stage entry points ---------------------------------------------------------------------------
int vh_upload_frame(vh_engine* e, const float* depth, const uint8_t* rgb) {
  if (!e || !depth) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  const size_t npx = (size_t)e->P.width * e->P.height;
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(e->d_depth[0], depth, npx * sizeof(float), cudaMemcpyHostToDevice));
  if (rgb) CK(cudaMemcpy(e->d_rgb[0], rgb, npx * 3, cudaMemcpyHostToDevice));
  e->cur_depth = e->d_depth[0];
  e->cur_rgb = (rgb && e->S.use_color) ? e->d_rgb[0] : nullptr;
  return VH_OK;
}

int vh_stage_allocate(vh_engine* e, const float* d_depth, const float* c2w) {
  if (!e || !c2w) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  if (d_depth) e->cur_depth = d_depth;
  setup_frame(e, c2w);
  CK(cudaMemsetAsync(e->D.counters, 0, sizeof(FrameCounters), e->stream));
  launch_alloc_visible(e->S, e->F, e->cur_depth, e->D, e->stream);
  int rc = enqueue_readback(e);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

int vh_stage_integrate(vh_engine* e, const float* d_depth, const uint8_t* d_rgb) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  if (d_depth) e->cur_depth = d_depth;
  if (d_rgb) e->cur_rgb = d_rgb;
  const int keep = e->S.use_color;
  e->S.use_color = e->cur_rgb ? keep : 0;
  CK(cudaMemsetAsync(&e->D.counters->voxel_updates, 0, sizeof(unsigned long long), e->stream));
  CK(cudaMemsetAsync(&e->D.counters->pad[0], 0, 2 * sizeof(unsigned long long), e->stream));     // verify mismatches, discarded blocks
  uint2* px = e->d_px[e->px_ring & 1]; e->px_ring++;
  launch_pack_frame(e->S, e->cur_depth, e->cur_rgb, px, e->D.tile_max, e->D.sched, nullptr, e->F.frame, e->stream);
  e->S.weight_bound = ++e->integrate_launches + e->weight_bound_bias;
  launch_cull_list(e->S, e->F, e->D, e->num_sms, e->stream);
  pick_integrate_parts(e);
  launch_integrate(e->S, e->F, px, e->cur_rgb != nullptr, e->D, e->num_sms, e->stream);
  e->S.use_color = keep;
  int rc = enqueue_readback(e);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

int vh_stage_marching_cubes(vh_engine* e) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  int rc = make_room(e);
  if (rc != VH_OK) return rc;
  CK(cudaMemsetAsync(&e->D.counters->triangles, 0, sizeof(unsigned long long), e->stream));
  launch_marching_cubes(e->S, e->F, e->D, e->D.visible, &e->D.counters->visible_count, 0, e->D.tri_offset, e->D.tri_count, e->num_sms, e->stream);
  rc = enqueue_readback(e);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

static int upload_keys(vh_engine* e, const int32_t* keys_xyz, int n) {
  if ((size_t)n > e->keys_tmp_cap) {
    cudaFree(e->d_keys_tmp);
    e->d_keys_tmp = nullptr;
    CK(cudaMalloc((void**)&e->d_keys_tmp, (size_t)n * sizeof(u64)));
    e->keys_tmp_cap = (size_t)n;
  }
  std::vector<u64> packed((size_t)n);
  for (int i = 0; i < n; i++) {
    const int x = keys_xyz[3 * i], y = keys_xyz[3 * i + 1], z = keys_xyz[3 * i + 2];
    if (!key_in_range(x, y, z)) return fail(VH_ERR_INVALID, "block coordinate (%d,%d,%d) outside [-2^20, 2^20)", x, y, z);
    packed[i] = pack_key(x, y, z);
  }
  CK(cudaMemcpyAsync(e->d_keys_tmp, packed.data(), (size_t)n * sizeof(u64), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return VH_OK;
}

int vh_set_visible(vh_engine* e, const int32_t* keys_xyz, int n, const float* c2w) {
  if (!e || !c2w || (n > 0 && !keys_xyz)) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  if (n > e->D.list_cap) return fail(VH_ERR_INVALID, "visible list of %d blocks exceeds capacity %d", n, e->D.list_cap);
  setup_frame(e, c2w);
  CK(cudaMemsetAsync(e->D.counters, 0, sizeof(FrameCounters), e->stream));
  if (n > 0) {
    int rc = upload_keys(e, keys_xyz, n);
    if (rc != VH_OK) return rc;
    launch_set_visible(e->D, e->d_keys_tmp, n, e->F.frame, e->stream);
  }
  int rc = enqueue_readback(e);
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return finish_sync(e);
}

int vh_get_stats(vh_engine* e, vh_stats* out) {
  if (!e || !out) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  memset(out, 0, sizeof(*out));
  out->frames = e->frames;
  out->visible_blocks = (uint32_t)e->h_block->c.visible_count;
  out->allocated_blocks = (uint32_t)e->h_block->heap_counter;
  out->voxel_updates = e->h_block->c.voxel_updates;
  out->voxel_updates_total = e->h_block->updates_total;
  out->triangles = e->h_block->c.triangles;
  out->arena_triangles = e->h_block->arena_top;
  out->debug_mismatches = e->h_block->c.pad[0];
  out->arena_compactions = e->compactions;
  out->forced_syncs = e->forced_syncs;
  out->culled_blocks = e->h_block->c.pad[1];
  if (e->frames > 0) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]) == cudaSuccess) out->ms_upload = ms;
    if (cudaEventElapsedTime(&ms, e->ev[1], e->ev[5]) == cudaSuccess) out->ms_alloc = ms;
    if (cudaEventElapsedTime(&ms, e->ev[5], e->ev[2]) == cudaSuccess) out->ms_cull = ms;
    if (cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]) == cudaSuccess) out->ms_integrate = ms;
    if (cudaEventElapsedTime(&ms, e->ev[3], e->ev[4]) == cudaSuccess) out->ms_mc = ms;
    cudaGetLastError();
  }
  return VH_OK;
}

void* vh_stream(vh_engine* e) { return e ? (void*)e->stream : nullptr; }

// ---- inspection -------------------------------------------------------------------------------------
__global__ void entries_to_keys_kernel(const u64* __restrict__ table, const int* __restrict__ list, int n, u64* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = table[list[i]];
}

int vh_visible_keys(vh_engine* e, int32_t* out_xyz, int cap, int* n_out) {
  if (!e || !n_out) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  int n = 0;
  CK(cudaMemcpy(&n, &e->D.counters->visible_count, sizeof(int), cudaMemcpyDeviceToHost));
  n = std::min(n, e->D.list_cap);
  *n_out = n;
  if (!out_xyz || n == 0) return VH_OK;
  const int m = std::min(n, cap);
  u64* d_tmp = nullptr;
  CK(cudaMalloc((void**)&d_tmp, (size_t)m * sizeof(u64)));
  entries_to_keys_kernel<<<(m + 255) / 256, 256, 0, e->stream>>>(e->D.map.keys, e->D.visible, m, d_tmp);
  std::vector<u64> keys((size_t)m);
  cudaError_t ce = cudaMemcpyAsync(keys.data(), d_tmp, (size_t)m * sizeof(u64), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(d_tmp);
  CK(ce);
  for (int i = 0; i < m; i++) unpack_key(keys[i], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
  return VH_OK;
}

int vh_allocated_keys(vh_engine* e, int32_t* out_xyz, int cap, int* n_out) {
  if (!e || !n_out) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  int n = 0;
  CK(cudaMemcpy(&n, e->D.map.heap_counter, sizeof(int), cudaMemcpyDeviceToHost));
  n = std::min(n, e->P.pool_blocks);
  *n_out = n;
  if (!out_xyz || n == 0) return VH_OK;
  const int m = std::min(n, cap);
  std::vector<u64> keys((size_t)m);
  CK(cudaMemcpy(keys.data(), e->D.map.key_heap, (size_t)m * sizeof(u64), cudaMemcpyDeviceToHost));
  for (int i = 0; i < m; i++) unpack_key(keys[i], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
  return VH_OK;
}

int vh_download_blocks(vh_engine* e, const int32_t* keys_xyz, int n, float* sdf, float* weight, uint8_t* rgb, uint8_t* found) {
  if (!e || (n > 0 && !keys_xyz)) return fail(VH_ERR_INVALID, "null argument");
  if (n <= 0) return VH_OK;
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  const int CH = 16384;   // blocks per staging round (64 MB of planes)
  float *d_s = nullptr, *d_w = nullptr; uint8_t *d_c = nullptr, *d_f = nullptr;
  const int m0 = std::min(n, CH);
  CK(cudaMalloc((void**)&d_s, (size_t)m0 * BLOCK_VOX * sizeof(float)));
  CK(cudaMalloc((void**)&d_w, (size_t)m0 * BLOCK_VOX * sizeof(float)));
  CK(cudaMalloc((void**)&d_c, (size_t)m0 * BLOCK_VOX * 3));
  CK(cudaMalloc((void**)&d_f, (size_t)m0));
  int rc = VH_OK;
  for (int o = 0; o < n && rc == VH_OK; o += CH) {
    const int m = std::min(CH, n - o);
    rc = upload_keys(e, keys_xyz + 3 * (size_t)o, m);
    if (rc != VH_OK) break;
    launch_gather_blocks(e->D, e->d_keys_tmp, m, d_s, d_w, d_c, d_f, e->stream);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess && sdf) ce = cudaMemcpy(sdf + (size_t)o * BLOCK_VOX, d_s, (size_t)m * BLOCK_VOX * sizeof(float), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && weight) ce = cudaMemcpy(weight + (size_t)o * BLOCK_VOX, d_w, (size_t)m * BLOCK_VOX * sizeof(float), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && rgb) ce = cudaMemcpy(rgb + (size_t)o * BLOCK_VOX * 3, d_c, (size_t)m * BLOCK_VOX * 3, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && found) ce = cudaMemcpy(found + o, d_f, (size_t)m, cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) rc = fail(VH_ERR_CUDA, "CUDA Error: %s in vh_download_blocks", cudaGetErrorString(ce));
  }
  cudaFree(d_s); cudaFree(d_w); cudaFree(d_c); cudaFree(d_f);
  return rc;
}

int vh_voxel_checksum(vh_engine* e, double* sum_sdf, double* sum_w, uint64_t* n_observed, uint64_t* n_negative) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  CK(cudaSetDevice(e->P.device));
  double* d4 = nullptr;
  CK(cudaMalloc((void**)&d4, 4 * sizeof(double)));
  launch_checksum(e->D, d4, e->stream);
  double h4[4];
  cudaError_t ce = cudaMemcpyAsync(h4, d4, sizeof(h4), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(d4);
  CK(ce);
  if (sum_sdf) *sum_sdf = h4[0];
  if (sum_w) *sum_w = h4[1];
  if (n_observed) *n_observed = (uint64_t)h4[2];
  if (n_negative) *n_negative = (uint64_t)h4[3];
  return VH_OK;
}

// ---- mesh assembly ------------------------------------------------------------------------------------

static inline int floor_div(int a, int b) { return (int)floorf((float)a / (float)b); }   // block2chunk, tsdf.cu:256-260

// collect (key, offset, count) records, sorted in tsdf2mesh order: chunk x,y,z ascending, then block-in-chunk linear
int collect_blocks(vh_engine* e, int mode, MeshBlocks& mb) {
  DeviceView D = e->D;
  const int nb = e->P.pool_blocks;
  const unsigned long long* d_off = D.tri_offset; const int* d_cnt = D.tri_count;
  mb.arena = D.arena;
  if (mode == VH_MESH_FULL_MAP) {
    if (!e->d_full_list) {
      CK(cudaMalloc((void**)&e->d_full_list, (size_t)nb * sizeof(int)));
      CK(cudaMalloc((void**)&e->d_full_count, sizeof(int)));
      CK(cudaMalloc((void**)&e->d_full_off, (size_t)nb * sizeof(unsigned long long)));
      CK(cudaMalloc((void**)&e->d_full_cnt, (size_t)nb * sizeof(int)));
    }
    // marching cubes over every allocated block into a scratch arena; grow until it fits
    unsigned long long cap = std::max<unsigned long long>(1ull << 20, D.arena_cap / 4);
    unsigned long long* d_top = nullptr;   // scratch arena top + {engine_error, overflow_frame} for this pass
    CK(cudaMalloc((void**)&d_top, 2 * sizeof(unsigned long long)));
    int* d_err = reinterpret_cast<int*>(d_top + 1);
    for (;;) {
      CK(cudaMalloc((void**)&mb.tmp_arena, cap * sizeof(vh_triangle)));
      DeviceView V = D;
      V.arena = mb.tmp_arena; V.arena_cap = cap; V.arena_top = d_top; V.list_cap = nb;
      V.engine_error = d_err; V.overflow_frame = reinterpret_cast<uint32_t*>(d_err + 1);
      CK(cudaMemsetAsync(d_top, 0, 2 * sizeof(unsigned long long), e->stream));
      CK(cudaMemsetAsync(e->d_full_cnt, 0, (size_t)nb * sizeof(int), e->stream));
      launch_list_all_blocks(D, e->d_full_list, e->d_full_count, e->stream);
      launch_marching_cubes(e->S, e->F, V, e->d_full_list, e->d_full_count, 1, e->d_full_off, e->d_full_cnt, e->num_sms, e->stream);
      unsigned long long top = 0; int err = 0;
      CK(cudaMemcpyAsync(&top, d_top, sizeof(top), cudaMemcpyDeviceToHost, e->stream));
      CK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      if (!(err & 1)) break;
      cudaFree(mb.tmp_arena); mb.tmp_arena = nullptr;
      cap = std::max(cap * 2, top + (top >> 2));
    }
    cudaFree(d_top);
    d_off = e->d_full_off; d_cnt = e->d_full_cnt; mb.arena = mb.tmp_arena;
  }
  u64* r_key = nullptr; unsigned long long* r_off = nullptr; int* r_cnt = nullptr; int* r_n = nullptr;
  CK(cudaMalloc((void**)&r_key, (size_t)nb * sizeof(u64)));
  CK(cudaMalloc((void**)&r_off, (size_t)nb * sizeof(unsigned long long)));
  CK(cudaMalloc((void**)&r_cnt, (size_t)nb * sizeof(int)));
  CK(cudaMalloc((void**)&r_n, sizeof(int)));
  launch_block_records(D, d_off, d_cnt, r_key, r_off, r_cnt, r_n, e->stream);
  int n = 0;
  CK(cudaMemcpyAsync(&n, r_n, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  std::vector<u64> key((size_t)n); std::vector<unsigned long long> off((size_t)n); std::vector<int> cnt((size_t)n);
  if (n) {
    CK(cudaMemcpy(key.data(), r_key, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(off.data(), r_off, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cnt.data(), r_cnt, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  }
  cudaFree(r_key); cudaFree(r_off); cudaFree(r_cnt); cudaFree(r_n);
  // order: (chunk x, chunk y, chunk z, block x, block y, block z) — equal to chunk order then (lx*bpc+ly)*bpc+lz
  struct Ord { int c[3]; int b[3]; int idx; };
  std::vector<Ord> ord((size_t)n);
  const int bpc = e->P.blocks_per_chunk;
  for (int i = 0; i < n; i++) {
    unpack_key(key[i], ord[i].b[0], ord[i].b[1], ord[i].b[2]);
    for (int a = 0; a < 3; a++) ord[i].c[a] = floor_div(ord[i].b[a], bpc);
    ord[i].idx = i;
  }
  std::sort(ord.begin(), ord.end(), [](const Ord& p, const Ord& q) {
    for (int a = 0; a < 3; a++) if (p.c[a] != q.c[a]) return p.c[a] < q.c[a];
    for (int a = 0; a < 3; a++) if (p.b[a] != q.b[a]) return p.b[a] < q.b[a];
    return false;
  });
  mb.key.resize(n); mb.off.resize(n); mb.cnt.resize(n);
  for (int i = 0; i < n; i++) { mb.key[i] = key[ord[i].idx]; mb.off[i] = off[ord[i].idx]; mb.cnt[i] = cnt[ord[i].idx]; }
  return VH_OK;
}

}  // extern "C"

// copy the triangle ranges of mb's blocks, in mb's order, into one array of `total` triangles: on the host (out) and/or left on
// the device (*d_keep, to be cudaFree'd by the caller)
int gather_block_triangles(vh_engine* e, const MeshBlocks& mb, vh_triangle* out, unsigned long long total, vh_triangle** d_keep) {
  if (d_keep) *d_keep = nullptr;
  const int n = (int)mb.key.size();
  if (total == 0 || n == 0) return VH_OK;
  std::vector<unsigned long long> dst((size_t)n);
  unsigned long long acc = 0;
  for (int i = 0; i < n; i++) { dst[i] = acc; acc += (unsigned long long)mb.cnt[i]; }
  unsigned long long *d_src = nullptr, *d_dst = nullptr; int* d_cnt = nullptr; vh_triangle* d_out = nullptr;
  cudaError_t ce = cudaMalloc((void**)&d_src, (size_t)n * sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_dst, (size_t)n * sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_cnt, (size_t)n * sizeof(int));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&d_out, (size_t)total * sizeof(vh_triangle));
  if (ce == cudaSuccess) ce = cudaMemcpy(d_src, mb.off.data(), (size_t)n * sizeof(unsigned long long), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_dst, dst.data(), (size_t)n * sizeof(unsigned long long), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_cnt, mb.cnt.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    launch_gather_triangles(mb.arena, d_src, d_dst, d_cnt, n, d_out, e->stream);
    ce = cudaStreamSynchronize(e->stream);
  }
  if (ce == cudaSuccess && out) ce = cudaMemcpy(out, d_out, (size_t)total * sizeof(vh_triangle), cudaMemcpyDeviceToHost);
  cudaFree(d_src); cudaFree(d_dst); cudaFree(d_cnt);
  if (ce != cudaSuccess || !d_keep) cudaFree(d_out); else *d_keep = d_out;
  if (ce != cudaSuccess) return fail(VH_ERR_CUDA, "CUDA Error: %s while gathering triangles", cudaGetErrorString(ce));
  return VH_OK;
}

extern "C" {

static int extract_mesh_locked(vh_engine* e, int mode, std::vector<vh_triangle>* host_out, vh_triangle* out, uint64_t cap, uint64_t* n_out) {
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  int rc = finish_sync(e);
  if (rc != VH_OK) return rc;
  MeshBlocks mb;
  rc = collect_blocks(e, mode, mb);
  if (rc != VH_OK) { cudaFree(mb.tmp_arena); return rc; }
  unsigned long long total = 0;
  for (int c : mb.cnt) total += (unsigned long long)c;
  if (n_out) *n_out = total;
  if (host_out) { host_out->resize((size_t)total); out = host_out->data(); cap = total; }
  if (out && total > cap) {
    cudaFree(mb.tmp_arena);
    return fail(VH_ERR_INVALID, "output capacity %llu < %llu triangles", (unsigned long long)cap, total);
  }
  if (out) rc = gather_block_triangles(e, mb, out, total, nullptr);
  cudaFree(mb.tmp_arena);
  return rc;
}

int vh_extract_mesh(vh_engine* e, int mode, vh_triangle* out, uint64_t cap, uint64_t* n) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  if (mode != VH_MESH_REF_PERSISTENT && mode != VH_MESH_FULL_MAP) return fail(VH_ERR_INVALID, "unknown mesh mode %d", mode);
  std::lock_guard<std::mutex> lk(e->mtx);
  return extract_mesh_locked(e, mode, nullptr, out, cap, n);
}

// welded mesh of the whole map: ordered soup gathered on the device, welded there (vh_weld.cu), copied out once
static int welded_mesh_locked(vh_engine* e, int mode, std::vector<vh_vertex>& v, std::vector<int32_t>& f) {
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  int rc = finish_sync(e);
  if (rc != VH_OK) return rc;
  MeshBlocks mb;
  rc = collect_blocks(e, mode, mb);
  if (rc != VH_OK) { cudaFree(mb.tmp_arena); return rc; }
  unsigned long long total = 0;
  for (int c : mb.cnt) total += (unsigned long long)c;
  vh_triangle* d_soup = nullptr;
  rc = gather_block_triangles(e, mb, nullptr, total, &d_soup);
  cudaFree(mb.tmp_arena);
  if (rc == VH_OK) rc = weld_on_device(e, d_soup, total, v, f);
  cudaFree(d_soup);
  return rc;
}

int vh_weld_mesh(vh_engine* e, int mode, vh_vertex* verts, uint64_t vcap, uint64_t* nv, int32_t* faces, uint64_t fcap, uint64_t* nf) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  if (mode != VH_MESH_REF_PERSISTENT && mode != VH_MESH_FULL_MAP) return fail(VH_ERR_INVALID, "unknown mesh mode %d", mode);
  std::lock_guard<std::mutex> lk(e->mtx);
  std::vector<vh_vertex> v; std::vector<int32_t> f;
  int rc = welded_mesh_locked(e, mode, v, f);
  if (rc != VH_OK) return rc;
  if (nv) *nv = v.size();
  if (nf) *nf = f.size() / 3;
  if (verts) { if (vcap < v.size()) return fail(VH_ERR_INVALID, "vertex capacity too small"); memcpy(verts, v.data(), v.size() * sizeof(vh_vertex)); }
  if (faces) { if (fcap < f.size() / 3) return fail(VH_ERR_INVALID, "face capacity too small"); memcpy(faces, f.data(), f.size() * sizeof(int32_t)); }
  return VH_OK;
}

// ASCII: same header and default ostream number formatting as tsdf2mesh (tsdf.cu:1870-1885). Binary: the same elements as
// binary_little_endian 1.0 (15-byte vertices, 13-byte faces), exact floats and a fraction of the size.
static int save_ply(vh_engine* e, const char* path, int mode, bool binary) {
  if (!e || !path) return fail(VH_ERR_INVALID, "null argument");
  if (mode != VH_MESH_REF_PERSISTENT && mode != VH_MESH_FULL_MAP) return fail(VH_ERR_INVALID, "unknown mesh mode %d", mode);
  std::lock_guard<std::mutex> lk(e->mtx);
  std::vector<vh_vertex> v; std::vector<int32_t> f;
  int rc = welded_mesh_locked(e, mode, v, f);
  if (rc != VH_OK) return rc;
  std::ofstream ply(path, binary ? std::ios::binary : std::ios::out);
  if (!ply) return fail(VH_ERR_IO, "cannot open %s", path);
  ply << "ply\nformat " << (binary ? "binary_little_endian" : "ascii") << " 1.0\ncomment stanford bunny\nelement vertex " << v.size() << "\n";
  ply << "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n";
  ply << "element face " << f.size() / 3 << "\n";
  ply << "property list uchar int vertex_index\nend_header\n";
  if (binary) {
    std::vector<char> buf;
    buf.reserve(v.size() * 15 + f.size() / 3 * 13);
    for (const auto& p : v) { const char* q = reinterpret_cast<const char*>(&p); buf.insert(buf.end(), q, q + 15); }   // x y z r g b
    for (size_t t = 0; t < f.size() / 3; t++) { buf.push_back(3); const char* q = reinterpret_cast<const char*>(&f[3 * t]); buf.insert(buf.end(), q, q + 12); }
    ply.write(buf.data(), (std::streamsize)buf.size());
  } else {
    for (const auto& p : v) ply << p.x << " " << p.y << " " << p.z << " " << (int)p.r << " " << (int)p.g << " " << (int)p.b << "\n";
    for (size_t t = 0; t < f.size() / 3; t++) ply << "3 " << f[3 * t] << " " << f[3 * t + 1] << " " << f[3 * t + 2] << "\n";
  }
  ply.close();
  if (!ply) return fail(VH_ERR_IO, "write to %s failed", path);
  return VH_OK;
}
int vh_save_ply(vh_engine* e, const char* path, int mode) { return save_ply(e, path, mode, false); }
int vh_save_ply_binary(vh_engine* e, const char* path, int mode) { return save_ply(e, path, mode, true); }

int vh_host_alloc(void** p, size_t bytes) {
  if (!p) return fail(VH_ERR_INVALID, "null argument");
  CK(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
  return VH_OK;
}
int vh_host_free(void* p) { CK(cudaFreeHost(p)); return VH_OK; }

}  // extern "C"
