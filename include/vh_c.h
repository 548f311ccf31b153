/* vh_c.h — C ABI of the B200-native voxel-hashing TSDF engine (libvhsdf.so).
 *
 * This is the drop-in boundary for ONE path of E-BAO/Voxel-Hashing-SDF: the per-frame
 * block allocation -> TSDF integrate -> marching cubes behind
 *     ark::GpuTsdfGenerator::processFrame   (/root/reference/include/tsdf.cuh:610, src/tsdf.cu:1485-1598)
 * its mesh export
 *     ark::GpuTsdfGenerator::SavePLY        (/root/reference/include/tsdf.cuh:628, src/tsdf.cu:1697-1888)
 * and the spatial hash those run on
 *     vhashing::HashTable / HashTableBase   (/root/reference/include/vhashing.h:35-826).
 * The C++ classes of the same names in include/tsdf.cuh and include/vhashing.h of THIS repo are thin
 * wrappers over these entry points; INTEGRATION.md shows the binding a reference maintainer adds.
 *
 * Plain C types only: pointers, sizes, ints, floats. No torch/thrust/STL types cross this line.
 * Every function returns a vh_status; vh_last_error() gives the message of the calling thread's
 * last failure. There is no CPU fallback: vh_create fails with VH_ERR_NO_DEVICE without a GPU.
 */
#ifndef VH_C_H_
#define VH_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VH_API __attribute__((visibility("default")))

typedef enum vh_status {
  VH_OK = 0,
  VH_ERR_INVALID = 1,      /* bad argument / unsupported parameter value */
  VH_ERR_NO_DEVICE = 2,    /* no CUDA device, or not an sm_100 part */
  VH_ERR_CUDA = 3,         /* CUDA runtime failure (reference: cudaSafeCall throws "CUDA Error", safecall.cpp:9-28) */
  VH_ERR_TABLE_FULL = 4,   /* hash table has no free entry (reference: operator[] spins forever, vhashing.h:216-231) */
  VH_ERR_POOL_FULL = 5,    /* voxel-block pool exhausted (reference: "out of block memory", blockalloc.h:51) */
  VH_ERR_ARENA_FULL = 6,   /* triangle arena could not grow */
  VH_ERR_IO = 7,
  VH_ERR_NOT_FOUND = 8
} vh_status;

/* mesh extraction modes */
#define VH_MESH_REF_PERSISTENT 0 /* reference semantics: per block, triangles of the last frame that saw it (streamOutGPU2CPU, tsdf.cu:534-540) */
#define VH_MESH_FULL_MAP 1       /* marching cubes over every allocated block, neighbours = allocated blocks */

typedef struct vh_engine vh_engine;

/* Every compile-time constant of the reference is a runtime parameter here.
 * vh_default_params() fills the reference's values except voxels_per_block (8, fixed by the engine). */
typedef struct vh_params {
  int width, height;            /* depth image size                                  (ctor args, tsdf.cuh:606) */
  float fx, fy, cx, cy;         /* pinhole intrinsics                                (tsdf.cu:1261-1266)       */
  float min_depth;              /* 0.1                                               (tsdf.cu:1318)            */
  float max_depth;              /* MaxDepth                                          (scene0220_02.yaml:37)    */
  float vox_size;               /* Voxel.Size                                                                  */
  float trunc_margin;           /* Voxel.TruncMargin                                                           */
  int voxels_per_block;         /* must be 8 (reference macro VOXEL_PER_BLOCK 5,     tsdf.cuh:40)              */
  int blocks_per_chunk;         /* BLOCK_PER_CHUNK 8                                 (tsdf.cuh:41)             */
  int dda_stride;               /* DDA_STEP 10                                       (tsdf.cu:13)              */
  int max_ray_steps;            /* maxLoopIterCount 100                              (tsdf.cu:2156)            */
  float chunk_radius;           /* CHUNK_RADIUS 4.0                                  (tsdf.cuh:44)             */
  int max_chunk_num;            /* MAX_CHUNK_NUM 128 (world = +-64 chunks); 0 = unbounded (tsdf.cuh:43)        */
  int num_buckets;              /* hash shape as the reference spells it             (tsdf.cu:1488)            */
  int entries_per_bucket;       /* table capacity = next pow2 >= buckets*entries                               */
  int pool_blocks;              /* voxel-block pool size (reference heap: 400000)                              */
  int use_color;                /* 1: integrate rgb as the reference does; 0: skip colour planes              */
  int mc_per_frame;             /* 1: working-set marching cubes every frame (tsdf.cu:1544); 0: on demand      */
  int device;                   /* CUDA device ordinal                                                         */
  int shard_rank, shard_count;  /* multi-GPU: this engine owns blocks with owner(key) == shard_rank            */
  int shard_group;              /* ownership granularity, blocks per axis; 0 = blocks_per_chunk (8): neighbours
                                   mostly share an owner, so few marching-cubes halo reads cross NVLink          */
  int depth_tile_smem;          /* reserved (staging depth tiles in shared memory was measured not to pay: DESIGN.md 5.2) */
  uint64_t tri_arena_bytes;     /* triangle arena size (grows on demand); 0 = default 1 GiB. With mc_per_frame a
                                   second arena of the same size is held as the compaction target                */
} vh_params;

typedef struct vh_stats {
  uint64_t frames;              /* frames integrated so far */
  uint32_t visible_blocks;      /* |visible set| of the last frame (reference h_heapBlockCounter) */
  uint32_t allocated_blocks;    /* blocks ever allocated in the map */
  uint64_t voxel_updates;       /* voxels whose weight was incremented in the last frame */
  uint64_t voxel_updates_total; /* voxel updates since vh_create / vh_reset */
  uint64_t triangles;           /* valid triangles emitted for the last frame's working set */
  uint64_t arena_triangles;     /* triangles currently held in the arena (live + superseded) */
  float ms_upload, ms_alloc, ms_integrate, ms_mc;   /* CUDA-event times of the last frame's stages */
  uint64_t debug_mismatches;    /* with env VH_INTEGRATE_VERIFY=1: fast-path vs IEEE-path disagreements in the last frame (must be 0) */
  uint64_t arena_compactions;   /* times the triangle arena was compacted (superseded per-block meshes dropped) */
  uint64_t forced_syncs;        /* times an asynchronous call had to drain the stream to bound arena use */
  uint64_t culled_blocks;       /* visible blocks of the last frame discarded whole before integration (provably no update) */
  float ms_cull;                /* CUDA-event time of the pass that builds integrate's work list (whole-block discard); not part of ms_alloc / ms_integrate */
  float reserved_f;
} vh_stats;

/* vertex layout of the triangle soup: the reference's Vertex (tsdf.cuh:65-77), 16 bytes */
typedef struct vh_vertex { float x, y, z; uint8_t r, g, b, pad; } vh_vertex;
typedef struct vh_triangle { vh_vertex p[3]; } vh_triangle;   /* 48 bytes; only valid triangles are ever stored */

VH_API const char* vh_last_error(void);
VH_API const char* vh_version(void);
VH_API int vh_default_params(vh_params* p);

/* lifetime: replaces GpuTsdfGenerator ctor / Shutdown (tsdf.cu:1249-1372, :1621-1638) */
VH_API int vh_create(const vh_params* p, vh_engine** out);
VH_API int vh_destroy(vh_engine* e);
VH_API int vh_reset(vh_engine* e);                 /* drop the whole map, keep allocations */

/* --- the hot path: replaces GpuTsdfGenerator::processFrame (tsdf.cu:1485-1598) -------------------
 * depth: float32[height*width] metres row-major, 0 = invalid.  rgb: uint8[height*width*3] or NULL.
 * c2w: float32[16] row-major camera->world. Host pointers, borrowed for the duration of the call
 * (pinned or pageable). Synchronous like the reference: results are visible when it returns. */
VH_API int vh_integrate(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w);
/* Same work, asynchronous on the engine's stream: returns after enqueueing. The host buffers must stay
 * untouched until vh_sync() (or vh_wait_uploads()) returns. Uploads of frame k+1 overlap kernels of frame k. */
VH_API int vh_integrate_async(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w);
/* Depth as raw u16 samples (the reference's depth PNGs, SaveFrame.cpp:174-180): metres = (float)sample * depth_scale,
 * computed on the GPU exactly as frameLoad's convertTo + scale does; half the upload bytes of the f32 form. */
VH_API int vh_integrate_u16_async(vh_engine* e, const uint16_t* depth_u16, double depth_scale, const uint8_t* rgb, const float* c2w);
VH_API int vh_wait_uploads(vh_engine* e);
VH_API int vh_sync(vh_engine* e);
/* Inputs already resident in HBM (device pointers on the engine's device); c2w stays a host pointer. */
VH_API int vh_integrate_device(vh_engine* e, const float* d_depth, const uint8_t* d_rgb, const float* c2w);
/* stage-level entry points (what vh_integrate* chains); used by stage tests and benchmarks.
 * vh_upload_frame copies host images into the engine's own device buffers; the stage calls use those
 * when their device-pointer arguments are NULL. */
VH_API int vh_upload_frame(vh_engine* e, const float* depth, const uint8_t* rgb);
VH_API int vh_stage_allocate(vh_engine* e, const float* d_depth, const float* c2w);
VH_API int vh_stage_integrate(vh_engine* e, const float* d_depth, const uint8_t* d_rgb);
VH_API int vh_stage_marching_cubes(vh_engine* e);
/* replace the frame's visible list with caller-supplied block keys (allocating them if needed) */
VH_API int vh_set_visible(vh_engine* e, const int32_t* keys_xyz, int n, const float* c2w);

VH_API int vh_get_stats(vh_engine* e, vh_stats* out);
VH_API void* vh_stream(vh_engine* e);              /* the engine's cudaStream_t, for event timing by the caller */

/* --- inspection / export ---------------------------------------------------------------------- */
VH_API int vh_visible_keys(vh_engine* e, int32_t* out_xyz, int cap, int* n);        /* last frame, unordered */
VH_API int vh_allocated_keys(vh_engine* e, int32_t* out_xyz, int cap, int* n);      /* key_heap order (insertion) */
/* voxel planes of n blocks: sdf/weight float32[n*512], rgb uint8[n*512*3] (any may be NULL); found uint8[n] */
VH_API int vh_download_blocks(vh_engine* e, const int32_t* keys_xyz, int n, float* sdf, float* weight, uint8_t* rgb, uint8_t* found);
/* checksums over all allocated voxels (sum sdf, sum weight in double; observed and negative counts) */
VH_API int vh_voxel_checksum(vh_engine* e, double* sum_sdf, double* sum_w, uint64_t* n_observed, uint64_t* n_negative);

/* Triangle soup in the order tsdf2mesh walks it (tsdf.cu:1786-1806), voxel-index units.
 * Call with out == NULL to get the count. replaces the dense h_chunks[].tri_ arrays. */
VH_API int vh_extract_mesh(vh_engine* e, int mode, vh_triangle* out, uint64_t cap, uint64_t* n);
/* ASCII PLY identical in structure to tsdf2mesh's (vertex dedupe on exact xyz, xyz * vox_size) */
VH_API int vh_save_ply(vh_engine* e, const char* path, int mode);
/* the same mesh as binary_little_endian PLY: exact floats, ~4x smaller, no text formatting on the way */
VH_API int vh_save_ply_binary(vh_engine* e, const char* path, int mode);
/* welded mesh: unique vertices (scaled by vox_size) + faces, as SavePLY would write them; the dedupe runs on the GPU
 * (two stable radix sorts) and reproduces tsdf2mesh's numbering: ids in order of first appearance, first colour wins */
VH_API int vh_weld_mesh(vh_engine* e, int mode, vh_vertex* verts, uint64_t vcap, uint64_t* nv, int32_t* faces, uint64_t fcap, uint64_t* nf);

/* --- multi-GPU: one map sharded over several B200s, one engine per process and GPU ------------------------------
 * (new: the reference is single-GPU.) Create every engine with the same vh_params except device and shard_rank
 * (shard_count = number of GPUs); a block belongs to shard
 * vh_owner_of_block(x, y, z, shard_count, shard_group): ownership is hashed per cube of shard_group^3 blocks. Rank 0 obtains an
 * NCCL id (vh_shard_unique_id) and passes it to the other processes by any means; every rank then calls
 * vh_shard_connect (collective): NCCL communicator + CUDA-IPC mappings of the peers' tables, voxel planes, key inboxes, flag words and
 * rank 0's frame ring, which the kernels address directly over NVLink.
 * vh_integrate_sharded (collective, same order on every rank): rank 0 puts {pose, depth, rgb} into its frame ring and every other GPU's
 * copy engine pulls it over NVLink one frame ahead (env VH_SHARD_BCAST=nccl: ncclBroadcast instead); the rays are split across the
 * GPUs and every block key travels to its owner's inbox; every GPU allocates, integrates and meshes its own blocks, reading
 * marching-cubes halos out of the owner's memory. depth / rgb are read on rank 0 only (vh_integrate_sharded_device: device pointers);
 * c2w may be NULL on the other ranks (they then take the pose out of the frame at the price of a stream sync per frame).
 * Asynchronous like vh_integrate_async: call vh_sync before reading results. vh_reset of a sharded engine is collective.
 * vh_shard_gather_mesh (collective): the whole map's triangle soup on rank 0, merged in tsdf2mesh order, equal to the
 * single-GPU result; other ranks get *n = 0. vh_shard_stats (collective): group-wide sums of the last frame's counters. */
#define VH_NCCL_ID_BYTES 128
VH_API int vh_owner_of_block(int x, int y, int z, int shard_count, int shard_group);
VH_API int vh_shard_unique_id(uint8_t id[VH_NCCL_ID_BYTES]);
VH_API int vh_shard_connect(vh_engine* e, const uint8_t id[VH_NCCL_ID_BYTES]);
VH_API int vh_integrate_sharded(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w);
VH_API int vh_integrate_sharded_device(vh_engine* e, const float* d_depth, const uint8_t* d_rgb, const float* c2w);   /* frame resident in rank 0's HBM */
VH_API int vh_shard_barrier(vh_engine* e);        /* collective, stream-ordered: every GPU has finished what was enqueued before it */
VH_API int vh_shard_gather_mesh(vh_engine* e, int mode, vh_triangle* out, uint64_t cap, uint64_t* n);
VH_API int vh_shard_stats(vh_engine* e, vh_stats* sum);
/* host-only helper of the gather: merge per-shard block lists (each in mesh order) into the global mesh order */
VH_API int vh_mesh_order_merge(int n_parts, const int32_t* const* keys_xyz, const int* nblocks, int blocks_per_chunk, int32_t* out_part, int32_t* out_index);

/* ---- out-of-core tier (optional; call between frames, single-GPU maps) -------------------------------------------------
 * The reference keeps its map on the host and streams it through the GPU every frame (streamInCPU2GPU tsdf.cu:277-457,
 * streamOutGPU2CPU :469-596). Here the map is resident; these three calls move whole blocks for scenes beyond one GPU.
 * vh_far_blocks: allocated blocks whose chunk fails the reference's residency rule for pose c2w (chunk cube + chunk sphere
 *   around the frustum centre, tsdf.cu:166-187,300-312); *n = how many there are (may exceed cap), keys sorted.
 * vh_evict_blocks: like vh_download_blocks, then the blocks leave the map (entries, pool slots, stored triangles).
 * vh_upload_blocks: inserts the blocks with these voxels (an existing block is overwritten); rgb may be NULL (zeros).
 * The per-frame meshes of evicted blocks are dropped; blocks that come back have none until a frame sees them again
 * (VH_MESH_FULL_MAP re-meshes everything that is on the device). */
VH_API int vh_far_blocks(vh_engine* e, const float* c2w, int32_t* out_xyz, int cap, int* n);
/* the same rule for blocks that are in the caller's store, not on the device (host arithmetic only, no engine): out[i] = 1 if
 * block i belongs on the device for pose c2w — what streamInCPU2GPU would upload */
VH_API int vh_blocks_resident(const vh_params* p, const float* c2w, const int32_t* keys_xyz, int n, uint8_t* out);
VH_API int vh_evict_blocks(vh_engine* e, const int32_t* keys_xyz, int n, float* sdf, float* weight, uint8_t* rgb, uint8_t* found);
VH_API int vh_upload_blocks(vh_engine* e, const int32_t* keys_xyz, int n, const float* sdf, const float* weight, const uint8_t* rgb);

/* pinned host memory helpers for callers that want DMA-able frame buffers */
VH_API int vh_host_alloc(void** p, size_t bytes);
VH_API int vh_host_free(void* p);

/* --- the spatial hash on its own: replaces vhashing::HashTable<int3,...> (vhashing.h:627-826) ------
 * Lock-free 64-bit-CAS open-addressing table of block coordinates -> value slots, with the
 * reference's insertion-ordered key_heap/heap_counter. Keys are int32 xyz triples in [-2^20, 2^20). */
typedef struct vh_map vh_map;
VH_API int vh_map_create(int num_buckets, int entries_per_bucket, int num_blocks, int device, vh_map** out);
VH_API int vh_map_destroy(vh_map* m);
/* bulk device-side operations over host key arrays (AllocKeys / find / erase, vhashing.h:531-603) */
VH_API int vh_map_insert(vh_map* m, const int32_t* keys_xyz, int n, int32_t* out_slots);  /* slot per key, duplicates share */
VH_API int vh_map_find(vh_map* m, const int32_t* keys_xyz, int n, int32_t* out_slots);    /* -1 when absent */
VH_API int vh_map_erase(vh_map* m, const int32_t* keys_xyz, int n, int32_t* out_erased);  /* 1 if the key was present */
VH_API int vh_map_size(vh_map* m, int* n);
VH_API int vh_map_keys(vh_map* m, int32_t* out_xyz, int cap, int* n);                     /* key_heap[0..heap_counter) */
/* raw device view for callers' own kernels (see include/vhashing.h of this repo) */
typedef struct vh_map_view {
  unsigned long long* keys; int32_t* slots; uint32_t capacity_mask;
  int32_t* free_list; int32_t* free_top; unsigned long long* key_heap; int32_t* heap_counter; int32_t* error_flag;
  int32_t num_blocks;
} vh_map_view;
VH_API int vh_map_get_view(vh_map* m, vh_map_view* out);

#ifdef __cplusplus
}
#endif
#endif /* VH_C_H_ */
