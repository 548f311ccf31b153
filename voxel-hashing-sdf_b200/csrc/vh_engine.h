// vh_engine.h — internal layout of the engine (host + device views). Not part of the C ABI.
//
// HBM layout (all resident for the life of the engine; nothing is streamed to the host on the hot path,
// unlike the reference's per-frame chunk stream-in/out, /root/reference/src/tsdf.cu:277-457, :469-596):
//   map.keys   u64 [capacity]        packed block coordinate per entry        8 B/entry
//   map.slots  i32 [capacity]        pool slot of the entry's block           4 B/entry
//   stamps     u32 [capacity]        frame number of the last frame that saw the block
//   sdf        f32 [pool_blocks*512] TSDF plane,   voxel index (x*8+y)*8+z    2 KB/block
//   wgt        f32 [pool_blocks*512] weight plane                             2 KB/block
//   rgb        u8x4[pool_blocks*512] colour plane (only when use_color)       2 KB/block
//   neg_count  i32 [pool_blocks]     voxels with sdf < 0 in the block: lets marching cubes skip one-sign neighbourhoods
//   visible    i32 [list_cap]        entry indices of the frame's visible set (compacted)
//   tri_arena  48 B triangles, bump-allocated per block per frame; tri_offset/tri_count per pool slot
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vh_c.h"
#include "../../include/vh_map.cuh"

namespace vh {

constexpr int VPB = 8;            // voxels per block edge (the reference's VOXEL_PER_BLOCK is 5; BASELINE configs use 8)
constexpr int BLOCK_VOX = 512;

struct StaticParams {
  int W, H;
  float fx, fy, cx, cy;
  float min_depth, max_depth;
  float vox_size, trunc, block_size, chunk_size;
  float half_vox;                 // 0.5f * vox_size (tsdf.cu:2146)
  int stride, max_steps, bpc;
  int nrx, nry;                   // sampled rays per row / column (reference launch shape, tsdf.cu:2263-2264)
  int use_color;
  uint32_t shard_rank, shard_count;
  int shard_group;                // ownership granularity in blocks per axis
  float round_eps;                // distance from a .5 pixel tie below which integrate re-projects with IEEE divisions
  int verify;                     // debug: run fast and IEEE paths side by side and count disagreements
  uint32_t weight_bound;          // upper bound of any voxel weight after the coming integrate launch (= launches since reset)
  int integrate_cull;             // 1 (default): discard whole blocks behind everything seen in their footprint; 0: gate every voxel
  int integrate_parts;            // staged kernel: work items per block for the coming launch: 1 = whole blocks, 2 = x-halves (set per frame by the host)
  int integrate_ctas_per_sm;      // resident CTAs per SM (tuning, VH_INTEGRATE_CTAS; 0 = the kernel's default: direct 3 [or 4], staged halves 5 [or 4])
  uint32_t byte_bias;             // 0x4B000000 (bits of 2^23), read from the parameter block by integrate_kernel_r1's byte -> float permutes
  int integrate_rev;              // 1: integrate_kernel_direct; 2: integrate_kernel_staged (planes through shared memory by bulk async copies)
  int alloc_rev;                  // 0: alloc_visible_kernel (one kernel, sequential DDA per ray); 2: ray_keys_kernel + insert_keys_kernel (DDA as a merge, keys routed to their owner)
  int mc_rev;                     // marching-cubes mesh kernel: 1 = the block's colour tile is staged in shared memory by cp.async while pass 1 runs (VH_MC_COLOR_TILE)
  // sizeof(StaticParams) stays a multiple of 16: the FrameParams that follows it in every kernel's parameter block keeps its
  // 16-byte alignment, so its pose is still fetched with 128-bit constant loads (add fields four ints at a time)
};

struct FrameParams {
  float c2w[16];
  float fc[3];                    // frustum centre (tsdf.cu:154-161)
  int cstart[3], cend[3];         // candidate chunk cube (tsdf.cu:304-312)
  float chunk_test_radius;        // 0.5*CHUNK_RADIUS*sqrt(3)*1.1 (tsdf.cu:172)
  uint32_t frame;                 // 1-based frame stamp
};

// per-frame device counters (one 64-byte block, reset at the start of every frame by pack_frame_kernel)
struct FrameCounters {
  int visible_count;
  uint32_t frame;                 // frame stamp these counters belong to
  unsigned long long voxel_updates;
  unsigned long long triangles;
  unsigned long long pad[5];
};

// Everything the host reads back after a frame, contiguous in HBM so that ONE small D2H copy fetches it
// (the reference reads its heap counter with a 1x1 kernel and four memcpys, tsdf.cu:2318-2337).
struct DeviceStatus {
  FrameCounters c;
  int map_error;                  // MapError bits, sticky
  int heap_counter;               // number of allocated blocks (key_heap length)
  int engine_error;               // sticky: 1 = triangle arena overflow
  uint32_t overflow_frame;        // first frame whose marching cubes ran out of arena (0 = none); must follow engine_error
  unsigned long long arena_top;   // triangles reserved in the arena
  unsigned long long updates_total;   // voxel updates since the last reset
  unsigned long long pad[4];
};

// What marching cubes needs from a shard of a multi-GPU map: its table and its voxel planes. Own pointers for the own
// rank, CUDA-IPC mappings of the peers' allocations otherwise (read over NVLink by the kernel itself).
constexpr int MAX_SHARDS = 8;
struct PeerView {
  const u64* keys; const int* slots; const uint32_t* stamps; const int* neg_count; const float* sdf; const uchar4* rgb;
  u64* inbox; int* inbox_count;   // the shard's key inbox [2][inbox_cap] and its two fill counters (written by every peer's ray_keys_kernel)
  uint32_t mask; uint32_t pad;
};
struct PeerTable { PeerView v[MAX_SHARDS]; };

// marching-cubes work queue: blocks that passed the neighbourhood sign filter
struct McWork {
  int bx, by, bz, slot;
  int nb[8];                      // pool slots of the 8 corner blocks (bit0 +x, bit1 +y, bit2 +z), -1 = absent
  unsigned char owner[8];         // multi-GPU: shard holding each corner block
  unsigned present;               // nb[i] >= 0 as bits
  unsigned pad;
};
struct McQueueCtl { int count; int head; };

struct DeviceView {
  MapView map;
  uint32_t* stamps;
  float* sdf;
  float* wgt;
  uchar4* rgb;
  int* sched;                     // [9 * 32] work counters of the integrate kernel's block scheduler + [8 * 32] = length of its work list (zeroed by pack_frame_kernel)
  uint4* work;                    // [list_cap] integrate work list {key lo, key hi, slot, position in `visible`}: visible blocks that survive the whole-block discard
  float* tile_max;                // [ceil(H/16) * ceil(W/16)] maximum depth per 16x16-pixel tile of the current frame
  int* neg_count;                 // [pool_blocks] number of voxels with sdf < 0 per block, kept current by integrate_kernel
  int* visible;
  int list_cap;
  u64* inbox;                     // [2][inbox_cap] block keys of the frame awaiting insertion, double-buffered by frame parity (ray_keys_kernel -> insert_keys_kernel)
  int* inbox_count;               // [2] keys in each half; [2..3] = CTAs of insert_keys_kernel that have finished (the last one empties the half)
  int* inbox_done;
  int inbox_cap;
  FrameCounters* counters;
  // triangle store
  vh_triangle* arena;
  unsigned long long* arena_top;  // triangles used
  unsigned long long arena_cap;   // triangles
  unsigned long long* tri_offset; // [pool_blocks]
  int* tri_count;                 // [pool_blocks]
  int* engine_error;              // sticky: 1 = arena overflow this frame
  uint32_t* overflow_frame;       // first frame that overflowed (0 = none)
  unsigned long long* updates_total;   // running total of voxel updates (status block)
  const PeerTable* peers;         // multi-GPU: every shard's view (nullptr on a single GPU)
  McWork* mc_queue;               // [list_cap] marching-cubes work queue
  McQueueCtl* mc_ctl;             // [2] queue control, alternating per launch
  int* mc_parity;                 // host-side toggle selecting the control slot of the next launch
};

// kernels (defined in the .cu files)
void launch_alloc_visible(const StaticParams& S, const FrameParams& F, const float* d_depth, const DeviceView& D, cudaStream_t st);
bool alloc_uses_inbox(const StaticParams& S);
void launch_ray_keys(const StaticParams& S, const FrameParams& F, const float* d_depth, const DeviceView& D, int num_sms, cudaStream_t st);
void launch_insert_keys(const StaticParams& S, const FrameParams& F, const DeviceView& D, int num_sms, cudaStream_t st);
void launch_pack_frame(const StaticParams& S, const float* d_depth, const uint8_t* d_rgb, uint2* d_out, float* d_tile_max, int* d_sched,
                       FrameCounters* reset_counters, uint32_t frame, cudaStream_t st, int stamp_only = 0);
void launch_cull_list(const StaticParams& S, const FrameParams& F, const DeviceView& D, int num_sms, cudaStream_t st);
void launch_integrate(const StaticParams& S, const FrameParams& F, const uint2* d_frame_px, bool color, const DeviceView& D, int num_sms,
                      cudaStream_t st);
void launch_marching_cubes(const StaticParams& S, const FrameParams& F, const DeviceView& D, const int* list, const int* list_count,
                           int full_map, unsigned long long* out_offset, int* out_count, int num_sms, cudaStream_t st);
void launch_set_visible(const DeviceView& D, const unsigned long long* d_keys, int n, uint32_t frame, cudaStream_t st);
void launch_list_all_blocks(const DeviceView& D, int* list, int* list_count, cudaStream_t st);
void launch_gather_blocks(const DeviceView& D, const unsigned long long* d_keys, int n, float* sdf, float* wgt, uint8_t* rgb, uint8_t* found,
                          cudaStream_t st);
void launch_checksum(const DeviceView& D, double* d_out4, cudaStream_t st);
void launch_block_records(const DeviceView& D, const unsigned long long* offsets, const int* counts, unsigned long long* rec_key,
                          unsigned long long* rec_off, int* rec_cnt, int* n_out, cudaStream_t st);
void launch_gather_triangles(const vh_triangle* arena, const unsigned long long* src_off, const unsigned long long* dst_off, const int* cnt, int n,
                             vh_triangle* out, cudaStream_t st);
void upload_mc_tables();

}  // namespace vh
