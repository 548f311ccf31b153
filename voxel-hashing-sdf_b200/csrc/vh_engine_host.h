// vh_engine_host.h — host-side state of an engine, shared by vh_engine.cu (single-GPU path) and vh_shard.cu (multi-GPU
// path). Internal: nothing here is part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "vh_engine.h"

using namespace vh;

int fail(int code, const char* fmt, ...);      // records the calling thread's error message, returns code
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) return fail(VH_ERR_CUDA, "CUDA Error: %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #call); \
  } while (0)

struct vh_shard_state;                          // NCCL communicator, peer mappings (vh_shard.cu)

struct vh_engine {
  vh_params P;
  StaticParams S;
  FrameParams F;
  DeviceView D;
  int num_sms = 148;
  uint32_t capacity = 0;
  cudaStream_t stream = nullptr, upload = nullptr;
  float* d_depth[2] = {nullptr, nullptr};
  uint16_t* d_depth16[2] = {nullptr, nullptr};   // staging of u16 depth frames (vh_integrate_u16_async), allocated on first use
  uint8_t* d_rgb[2] = {nullptr, nullptr};
  uint2* d_px[2] = {nullptr, nullptr};     // packed {depth, rgb} records the integrate kernel reads
  int px_ring = 0;
  cudaEvent_t ev_uploaded[2] = {nullptr, nullptr}, ev_rgb[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  bool buf_used[2] = {false, false};
  cudaEvent_t ev_mapped_read = nullptr;  // recorded on the compute stream behind a ray pass that read the caller's pinned depth buffer in place
  bool mapped_read_pending = false;
  int ring = 0;
  const float* cur_depth = nullptr;     // device pointers the stage calls operate on
  const uint8_t* cur_rgb = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // frame start, inputs ready, integrate start, integrate end, frame end, work-list start
  // pinned read-back block: counters of the last frame + flags
  typedef DeviceStatus HostBlock;
  HostBlock* h_block = nullptr;
  DeviceStatus* h_block_dev = nullptr;  // device-side address of h_block when the status is published by a kernel (VH_STATUS_PUBLISH=1)
  DeviceStatus* d_status = nullptr;     // counters, error flags, heap counter, arena top: one block, one read-back copy
  uint64_t frames = 0, updates_total = 0;
  uint64_t max_tris_per_frame = 0, known_arena_top = 0;
  int frames_in_flight = 0;
  // full-map extraction scratch
  int* d_full_list = nullptr; int* d_full_count = nullptr; unsigned long long* d_full_off = nullptr; int* d_full_cnt = nullptr;
  u64* d_keys_tmp = nullptr; size_t keys_tmp_cap = 0;
  // arena compaction: spare arena (ping-pong) and scan scratch
  vh_triangle* arena_spare = nullptr;
  unsigned long long *d_scan_in = nullptr, *d_scan_out = nullptr; void* d_scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
  uint64_t compactions = 0, forced_syncs = 0;
  uint32_t integrate_launches = 0;      // since the last reset: bounds every voxel weight
  uint32_t weight_bound_bias = 0;       // added to the weight bound: forces the general colour path (env VH_INTEGRATE_EXACT_COLOR=1, or non-integer uploaded weights)
  uint32_t weight_bound_env = 0;        // the environment's share of it (survives vh_reset)
  int integrate_parts_forced = 0;       // env VH_INTEGRATE_PARTS (1 or 2); 0 = chosen per frame from the previous frame's work list (pick_integrate_parts)
  int mc_parity = 0;                    // which McQueueCtl slot the next marching-cubes launch uses
  uint32_t tombstones = 0;              // table entries released by vh_evict_blocks since the last rebuild (vh_stream.cu)
  // multi-GPU (vh_shard.cu)
  vh_shard_state* shard = nullptr;
  PeerTable* d_peers = nullptr;
  std::mutex mtx;
};

// mesh assembly records: (key, arena offset, count) per block, sorted in tsdf2mesh order
struct MeshBlocks {
  std::vector<u64> key; std::vector<unsigned long long> off; std::vector<int> cnt;
  const vh_triangle* arena = nullptr; vh_triangle* tmp_arena = nullptr;
};

void setup_frame(vh_engine* e, const float* c2w);
void shard_release(vh_engine* e);               // vh_shard.cu: called by vh_destroy
int shard_barrier(vh_engine* e);                // vh_shard.cu: stream-ordered barrier across the GPUs of a sharded map (enqueue only)
int gather_block_triangles(vh_engine* e, const MeshBlocks& mb, vh_triangle* out, unsigned long long total, vh_triangle** d_keep);   // ordered soup of mb's blocks
int weld_on_device(vh_engine* e, const vh_triangle* d_soup, unsigned long long T, std::vector<vh_vertex>& verts, std::vector<int32_t>& faces);   // vh_weld.cu
extern "C" {
void pick_integrate_parts(vh_engine* e);
int enqueue_stages(vh_engine* e, bool do_alloc, cudaEvent_t depth_ready, cudaEvent_t rgb_ready, const float* host_depth_mapped);
int enqueue_readback(vh_engine* e);
int finish_sync(vh_engine* e);
int make_room(vh_engine* e);
int collect_blocks(vh_engine* e, int mode, MeshBlocks& mb);
}
