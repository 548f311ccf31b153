// vh_shard.cu — one map sharded over several B200s (BASELINE config 4): one engine per process and GPU, block
// ownership by a hash of the block coordinate, every GPU allocates, integrates and meshes only its own blocks, meshes
// gathered on rank 0. The reference has no multi-GPU path (one global pair of tables, no cudaSetDevice); this is new.
//   * vh_shard_connect: NCCL communicator (ncclCommInitRank) + an all-gather of CUDA-IPC handles: every process maps its peers'
//     tables, voxel planes, key inboxes, flag words and rank 0's frame ring (cudaIpcOpenMemHandle) — NVLink peer memory that
//     the kernels address directly.
//   * vh_integrate_sharded / vh_integrate_sharded_device, per frame (collective, same order on every rank):
//       rank 0 puts {pose, depth, rgb} into a slot of its frame ring (H2D from the caller's host buffers on the upload stream,
//       double-buffered, or one device copy) and raises a sequence flag in every peer's memory;
//       every other GPU's upload stream waits for the flag and PULLS the frame out of rank 0's ring over NVLink with one copy-engine
//       transfer into its own ring, one frame ahead of its compute stream: no collective launch, no SM time (VH_SHARD_BCAST=nccl
//       selects a per-frame ncclBroadcast instead, kept as the baseline);
//       ray_keys_kernel: the rays are SPLIT across the GPUs; every key goes to its owner's inbox with NVLink stores;
//       frame barrier; insert_keys_kernel on the own inbox; work list; integrate of the own blocks;
//       second barrier; marching cubes — neighbour blocks are looked up in the table of the GPU their key hashes to and
//       their sdf / colour halos are read from that GPU's planes, inside the kernels.
//     The barrier that keeps a GPU from integrating frame f+1 while a peer still meshes frame f is the first barrier of frame f+1.
//   * vh_shard_gather_mesh: every rank's ordered per-block triangle ranges go to rank 0 (ncclSend/ncclRecv), which
//     merges them by the reference's mesh order (tsdf2mesh, tsdf.cu:1786-1806) — identical to the one-GPU soup.
// NCCL is bound at run time (dlopen/dlsym) so that a process that already carries an NCCL (e.g. the one PyTorch
// bundles) shares it and the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "vh_engine_host.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mtx;

int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mtx);
  if (g_nccl.lib) return VH_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // whatever the process already carries
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(VH_ERR_INVALID, "NCCL not found (libnccl.so.2): %s", dlerror());
#define VH_SYM(field, name)                                                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                               \
  if (!g_nccl.field) return fail(VH_ERR_INVALID, "NCCL symbol %s missing", name)
  VH_SYM(GetUniqueId, "ncclGetUniqueId"); VH_SYM(CommInitRank, "ncclCommInitRank"); VH_SYM(CommDestroy, "ncclCommDestroy");
  VH_SYM(Broadcast, "ncclBroadcast"); VH_SYM(AllReduce, "ncclAllReduce"); VH_SYM(AllGather, "ncclAllGather");
  VH_SYM(Send, "ncclSend"); VH_SYM(Recv, "ncclRecv"); VH_SYM(GroupStart, "ncclGroupStart"); VH_SYM(GroupEnd, "ncclGroupEnd");
  VH_SYM(GetErrorString, "ncclGetErrorString");
#undef VH_SYM
  g_nccl.lib = h;
  return VH_OK;
}

#define NK(call)                                                                                          \
  do {                                                                                                    \
    ncclResult_t _r = (call);                                                                             \
    if (_r != ncclSuccess) return fail(VH_ERR_CUDA, "NCCL Error: %s at %s:%d (%s)", g_nccl.GetErrorString(_r), __FILE__, __LINE__, #call); \
  } while (0)

constexpr int N_SHARED = 10;    // keys, slots, stamps, neg_count, sdf, rgb, flag words, key inbox, inbox counters, frame ring
constexpr int FLAG_WORDS = 2 * MAX_SHARDS;      // [0, MAX_SHARDS): barrier arrival epochs; [MAX_SHARDS]: sequence number of the newest frame in rank 0's ring
struct ShardExport {
  cudaIpcMemHandle_t h[N_SHARED];
  uint32_t capacity; int pool_blocks; int has_rgb; int device;
};

}  // namespace

struct vh_shard_state {
  ncclComm_t comm = nullptr;
  int rank = 0, count = 1;
  void* mapped[MAX_SHARDS][N_SHARED] = {};
  uint8_t* d_frame = nullptr;         // frame ring: {c2w[16] f32 | depth f32[H*W] | rgb u8[H*W*3]} x 2 slots; rank 0's is mapped by every peer
  const uint8_t* frame_src = nullptr; // where this rank's kernels read frames from: rank 0's ring (pull) or the own ring (NCCL broadcast)
  bool pull = true;                   // frames are read out of rank 0's memory by the pack kernel (default); false: ncclBroadcast into the own ring
  cudaEvent_t released[2] = {nullptr, nullptr};   // rank 0: every GPU is past the frame barrier of the slot's previous frame (nobody reads it any more)
  uint64_t frame_seq = 0;             // frames put into the ring so far
  size_t frame_bytes = 0;
  uint8_t* h_frame = nullptr;         // pinned staging of the same layout (rank 0 packs the caller's buffers here)
  int ring = 0;
  cudaEvent_t consumed[2] = {nullptr, nullptr}, uploaded[2] = {nullptr, nullptr}, arrived[2] = {nullptr, nullptr}; bool used[2] = {false, false};
  int* d_token = nullptr;             // 1-int all-reduce (connect-time barrier)
  uint32_t* d_flags = nullptr;        // [FLAG_WORDS] arrival epochs written by the peers (and by this GPU) over NVLink, frame sequence flag
  uint32_t* peer_flags[MAX_SHARDS] = {};
  uint32_t epoch = 0;
  float* h_pose = nullptr;            // pinned: pose read back on ranks that were not given one
};

void shard_release(vh_engine* e) {
  vh_shard_state* s = e->shard;
  if (!s) return;
  if (e->stream && s->peer_flags[s->rank]) shard_barrier(e);      // no peer is still reading this GPU's memory (a dead peer costs the 4 s watchdog)
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->upload) cudaStreamSynchronize(e->upload);
  for (int r = 0; r < s->count; r++)
    for (int k = 0; k < N_SHARED; k++)
      if (r != s->rank && s->mapped[r][k]) cudaIpcCloseMemHandle(s->mapped[r][k]);
  if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
  cudaFree(s->d_frame); cudaFree(s->d_token); cudaFree(s->d_flags);
  if (s->h_frame) cudaFreeHost(s->h_frame);
  if (s->h_pose) cudaFreeHost(s->h_pose);
  for (int i = 0; i < 2; i++) {
    if (s->consumed[i]) cudaEventDestroy(s->consumed[i]);
    if (s->uploaded[i]) cudaEventDestroy(s->uploaded[i]);
    if (s->arrived[i]) cudaEventDestroy(s->arrived[i]);
    if (s->released[i]) cudaEventDestroy(s->released[i]);
  }
  cudaFree(e->d_peers); e->d_peers = nullptr; e->D.peers = nullptr;
  delete s;
  e->shard = nullptr;
}

// Stream-ordered barrier across the GPUs without a collective: every GPU stores its arrival epoch into every peer's
// flag array through the NVLink mappings and spins on its own array until all peers have arrived (~2 us, against
// ~25 us for a one-element NCCL all-reduce). Work enqueued after it on any GPU sees everything enqueued before it on
// every GPU (the peers' kernels have completed; remote reads are served by the owner's L2).
struct FlagPtrs { uint32_t* p[MAX_SHARDS]; };
__global__ void shard_barrier_kernel(FlagPtrs peers, uint32_t* __restrict__ mine, int rank, int n, uint32_t epoch, int* __restrict__ error) {
  const int t = threadIdx.x;
  if (t < n) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(peers.p[t] + rank) = epoch;
    const long long t0 = clock64();
    while ((int)(*reinterpret_cast<volatile uint32_t*>(mine + t) - epoch) < 0) {
      if (clock64() - t0 > 8000000000ll) { atomicOr(error, 8); break; }     // ~4 s: a peer died; never hang the GPU
    }
    __threadfence_system();
  }
}
int shard_barrier(vh_engine* e) {
  vh_shard_state* s = e->shard;
  FlagPtrs fp;
  for (int q = 0; q < MAX_SHARDS; q++) fp.p[q] = s->peer_flags[q];
  shard_barrier_kernel<<<1, 32, 0, e->stream>>>(fp, s->d_flags, s->rank, s->count, ++s->epoch, e->D.engine_error);
  return VH_OK;
}
// rank 0, on its upload stream behind the copies of frame `seq` into the ring: tell every GPU the frame is there
__global__ void frame_ready_kernel(FlagPtrs peers, int n, uint32_t seq) {
  const int t = threadIdx.x;
  if (t < n) { __threadfence_system(); *reinterpret_cast<volatile uint32_t*>(peers.p[t] + MAX_SHARDS) = seq; }
}
// every other GPU, on its compute stream ahead of the frame's first kernel
__global__ void frame_wait_kernel(const uint32_t* __restrict__ mine, uint32_t seq, int* __restrict__ error) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((int)(*reinterpret_cast<const volatile uint32_t*>(mine + MAX_SHARDS) - seq) < 0) {
      if (clock64() - t0 > 8000000000ll) { atomicOr(error, 8); break; }     // ~4 s: rank 0 died; never hang the GPU
    }
    __threadfence_system();
  }
}

static int nccl_barrier(vh_engine* e) {
  vh_shard_state* s = e->shard;
  NK(g_nccl.AllReduce(s->d_token, s->d_token, 1, ncclInt, ncclSum, s->comm, e->stream));
  return VH_OK;
}

extern "C" {

// which shard owns a block (host-side twin of owner_of_key in include/vh_map.cuh)
int vh_owner_of_block(int x, int y, int z, int shard_count, int shard_group) {
  if (shard_count <= 0 || shard_group <= 0 || !key_in_range(x, y, z)) return -1;
  return (int)owner_of_block(x, y, z, (uint32_t)shard_count, shard_group);
}

int vh_shard_unique_id(uint8_t id[VH_NCCL_ID_BYTES]) {
  if (!id) return fail(VH_ERR_INVALID, "null argument");
  int rc = load_nccl();
  if (rc != VH_OK) return rc;
  static_assert(VH_NCCL_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  ncclUniqueId u;
  NK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, VH_NCCL_ID_BYTES);
  return VH_OK;
}

int vh_shard_connect(vh_engine* e, const uint8_t id[VH_NCCL_ID_BYTES]) {
  if (!e || !id) return fail(VH_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mtx);
  if (e->shard) return fail(VH_ERR_INVALID, "engine already connected");
  const int n = e->P.shard_count, r = e->P.shard_rank;
  if (n < 2 || n > MAX_SHARDS) return fail(VH_ERR_INVALID, "shard_count must be 2..%d (got %d)", MAX_SHARDS, n);
  int rc = load_nccl();
  if (rc != VH_OK) return rc;
  CK(cudaSetDevice(e->P.device));
  vh_shard_state* s = new vh_shard_state;
  s->rank = r; s->count = n;
  e->shard = s;
  ncclUniqueId u;
  memcpy(u.internal, id, VH_NCCL_ID_BYTES);
  NK(g_nccl.CommInitRank(&s->comm, n, u, r));
  const size_t npx = (size_t)e->P.width * e->P.height;
  s->frame_bytes = (64 + npx * 4 + npx * 3 + 255) / 256 * 256;
  CK(cudaMalloc((void**)&s->d_frame, 2 * s->frame_bytes));
  CK(cudaHostAlloc((void**)&s->h_frame, 2 * s->frame_bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&s->h_pose, 2 * 16 * sizeof(float), cudaHostAllocDefault));
  CK(cudaMalloc((void**)&s->d_token, sizeof(int)));
  CK(cudaMemset(s->d_token, 0, sizeof(int)));
  CK(cudaMalloc((void**)&s->d_flags, FLAG_WORDS * sizeof(uint32_t)));
  CK(cudaMemset(s->d_flags, 0, FLAG_WORDS * sizeof(uint32_t)));
  { const char* v = getenv("VH_SHARD_BCAST"); s->pull = !(v && v[0] == 'n'); }      // VH_SHARD_BCAST=nccl: the ncclBroadcast route
  for (int i = 0; i < 2; i++) {
    CK(cudaEventCreateWithFlags(&s->consumed[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->uploaded[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->arrived[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->released[i], cudaEventDisableTiming));
  }

  // exchange IPC handles of what a peer's marching cubes needs to read
  ShardExport mine;
  memset(&mine, 0, sizeof(mine));
  void* ptrs[N_SHARED] = {e->D.map.keys, e->D.map.slots, e->D.stamps, e->D.neg_count, e->D.sdf, e->D.rgb, s->d_flags, e->D.inbox, e->D.inbox_count, s->d_frame};
  for (int k = 0; k < N_SHARED; k++)
    if (ptrs[k]) CK(cudaIpcGetMemHandle(&mine.h[k], ptrs[k]));
  mine.capacity = e->capacity; mine.pool_blocks = e->P.pool_blocks; mine.has_rgb = e->D.rgb ? 1 : 0; mine.device = e->P.device;
  ShardExport* d_all = nullptr;
  CK(cudaMalloc((void**)&d_all, sizeof(ShardExport) * n));
  CK(cudaMemcpyAsync(d_all + r, &mine, sizeof(mine), cudaMemcpyHostToDevice, e->stream));
  NK(g_nccl.AllGather(d_all + r, d_all, sizeof(ShardExport), ncclChar, s->comm, e->stream));
  std::vector<ShardExport> all((size_t)n);
  CK(cudaMemcpyAsync(all.data(), d_all, sizeof(ShardExport) * n, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(d_all);

  PeerTable pt;
  memset(&pt, 0, sizeof(pt));
  for (int q = 0; q < n; q++) {
    if (all[q].capacity != e->capacity || all[q].has_rgb != mine.has_rgb)
      return fail(VH_ERR_INVALID, "shard %d was created with different parameters (capacity %u vs %u)", q, all[q].capacity, e->capacity);
    void* m[N_SHARED];
    for (int k = 0; k < N_SHARED; k++) {
      if (q == r) m[k] = ptrs[k];
      else if (k == 5 && !all[q].has_rgb) m[k] = nullptr;
      else if (k == 9 && q != 0) m[k] = nullptr;                  // only rank 0's frame ring is read by its peers
      else {
        cudaError_t ce = cudaIpcOpenMemHandle(&m[k], all[q].h[k], cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) return fail(VH_ERR_CUDA, "CUDA Error: cannot map shard %d's memory (%s): the GPUs need peer access (NVLink/NVSwitch)", q, cudaGetErrorString(ce));
      }
      s->mapped[q][k] = m[k];
    }
    PeerView& v = pt.v[q];
    v.keys = (const u64*)m[0]; v.slots = (const int*)m[1]; v.stamps = (const uint32_t*)m[2]; v.neg_count = (const int*)m[3];
    v.sdf = (const float*)m[4]; v.rgb = (const uchar4*)m[5]; v.mask = e->capacity - 1;
    v.inbox = (u64*)m[7]; v.inbox_count = (int*)m[8];
    s->peer_flags[q] = (uint32_t*)m[6];
    if (q == 0) s->frame_src = s->pull ? (const uint8_t*)m[9] : s->d_frame;
  }
  CK(cudaMalloc((void**)&e->d_peers, sizeof(PeerTable)));
  CK(cudaMemcpy(e->d_peers, &pt, sizeof(pt), cudaMemcpyHostToDevice));
  e->D.peers = e->d_peers;
  rc = nccl_barrier(e);        // every rank has mapped its peers before anyone stores into a flag array
  if (rc != VH_OK) return rc;
  CK(cudaStreamSynchronize(e->stream));
  return VH_OK;
}

// One frame on every GPU of the group (collective: every rank calls it, in the same order). depth / rgb are read on
// rank 0 only (host pointers like vh_integrate, or device pointers for vh_integrate_sharded_device); c2w may be NULL on the
// other ranks, which then take the pose out of the frame (one small D2H + stream sync per frame on those ranks; pass the pose
// everywhere to stay asynchronous).
static int integrate_sharded_common(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w, bool host_inputs) {
  vh_shard_state* s = e->shard;
  if (!s) return fail(VH_ERR_INVALID, "vh_shard_connect has not been called");
  if (s->rank == 0 && (!depth || !c2w)) return fail(VH_ERR_INVALID, "rank 0 must supply depth and pose");
  CK(cudaSetDevice(e->P.device));
  int rc = make_room(e);
  if (rc != VH_OK) return rc;
  const size_t npx = (size_t)e->P.width * e->P.height;
  const int b = s->ring & 1; s->ring++;
  const uint32_t seq = (uint32_t)(++s->frame_seq);
  uint8_t* dbuf = s->d_frame + (size_t)b * s->frame_bytes;                       // own ring slot (rank 0: the slot everybody reads)
  const bool with_rgb = e->S.use_color != 0;                      // group-wide: every rank was created with the same flag
  const size_t bytes = 64 + npx * 4 + (with_rgb ? npx * 3 : 0);
  CK(cudaEventRecord(e->ev[0], e->stream));
  // Rank 0 fills the slot on its upload stream, double-buffered: frame k+1 arrives over PCIe while frame k is being integrated and
  // meshed. The slot's previous frame (k-1) is no longer read once every GPU is past frame k-1's barrier (`released`).
  cudaStream_t up = e->upload;
  if (s->used[b]) CK(cudaStreamWaitEvent(up, (s->pull && s->rank == 0) ? s->released[b] : s->consumed[b], 0));
  if (s->rank == 0) {
    uint8_t* hbuf = s->h_frame + (size_t)b * s->frame_bytes;
    if (s->used[b]) CK(cudaEventSynchronize(s->uploaded[b]));    // staging slot b's previous upload (two frames ago) has left the host
    memcpy(hbuf, c2w, 64);
    CK(cudaMemcpyAsync(dbuf, hbuf, 64, cudaMemcpyHostToDevice, up));
    if (!host_inputs) {
      CK(cudaMemcpyAsync(dbuf + 64, depth, npx * 4, cudaMemcpyDeviceToDevice, up));
      if (with_rgb) { if (rgb) CK(cudaMemcpyAsync(dbuf + 64 + npx * 4, rgb, npx * 3, cudaMemcpyDeviceToDevice, up)); else CK(cudaMemsetAsync(dbuf + 64 + npx * 4, 0, npx * 3, up)); }
    } else {
      // pinned caller buffers are copied straight into the ring; pageable ones go through the pinned staging slot
      cudaPointerAttributes pa;
      const bool pinned = cudaPointerGetAttributes(&pa, depth) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                          (!with_rgb || !rgb || (cudaPointerGetAttributes(&pa, rgb) == cudaSuccess && pa.type == cudaMemoryTypeHost));
      cudaGetLastError();
      if (pinned) {
        CK(cudaMemcpyAsync(dbuf + 64, depth, npx * 4, cudaMemcpyHostToDevice, up));
        if (with_rgb) { if (rgb) CK(cudaMemcpyAsync(dbuf + 64 + npx * 4, rgb, npx * 3, cudaMemcpyHostToDevice, up)); else CK(cudaMemsetAsync(dbuf + 64 + npx * 4, 0, npx * 3, up)); }
      } else {
        memcpy(hbuf + 64, depth, npx * 4);
        if (with_rgb) { if (rgb) memcpy(hbuf + 64 + npx * 4, rgb, npx * 3); else memset(hbuf + 64 + npx * 4, 0, npx * 3); }
        CK(cudaMemcpyAsync(dbuf + 64, hbuf + 64, bytes - 64, cudaMemcpyHostToDevice, up));
      }
    }
    CK(cudaEventRecord(s->uploaded[b], up));
    if (s->pull) {      // the frame is in the ring: raise the sequence flag in every GPU's memory
      FlagPtrs fp;
      for (int q = 0; q < MAX_SHARDS; q++) fp.p[q] = s->peer_flags[q];
      frame_ready_kernel<<<1, 32, 0, up>>>(fp, s->count, seq);
    }
  }
  // Every other GPU fetches the frame on ITS upload stream — a spin on the sequence flag, then one copy-engine transfer out of rank 0's
  // ring over NVLink into the own ring — so frame k+1 arrives while frame k is being integrated and meshed: no collective launch, no
  // SM time, nothing on the compute stream's critical path (kernels reading rank 0's memory in place were measured at 8 GPUs:
  // the pack kernel's small loads over NVLink cost ~50 us per frame, profiles/r02g). VH_SHARD_BCAST=nccl: ncclBroadcast instead.
  const uint8_t* src = dbuf;
  if (s->pull) {
    if (s->rank != 0) {
      frame_wait_kernel<<<1, 32, 0, up>>>(s->d_flags, seq, e->D.engine_error);
      CK(cudaMemcpyAsync(dbuf, s->frame_src + (size_t)b * s->frame_bytes, bytes, cudaMemcpyDefault, up));
    }
  } else {
    NK(g_nccl.Broadcast(dbuf, dbuf, bytes, ncclChar, 0, s->comm, up));
  }
  CK(cudaEventRecord(s->arrived[b], up));
  CK(cudaStreamWaitEvent(e->stream, s->arrived[b], 0));
  CK(cudaEventRecord(e->ev[1], e->stream));
  float pose[16];
  if (c2w) memcpy(pose, c2w, sizeof(pose));
  else {
    CK(cudaMemcpyAsync(s->h_pose, src, 64, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memcpy(pose, s->h_pose, sizeof(pose));
  }
  e->cur_depth = reinterpret_cast<const float*>(src + 64);
  e->cur_rgb = with_rgb ? src + 64 + npx * 4 : nullptr;
  setup_frame(e, pose);
  DeviceView& D = e->D;
  uint2* px = e->d_px[e->px_ring & 1]; e->px_ring++;
  static const int dbg = getenv("VH_SHARD_DEBUG") ? atoi(getenv("VH_SHARD_DEBUG")) : 0;   // timing experiments only: 1 = no barriers, 2 = local-only MC
  launch_pack_frame(e->S, e->cur_depth, e->cur_rgb, px, D.tile_max, D.sched, D.counters, e->F.frame, e->stream);
  if (alloc_uses_inbox(e->S)) {
    // the rays are split across the GPUs, their keys travel to the owners' inboxes; once every GPU is past the barrier all keys
    // of the frame have arrived (and every GPU has finished meshing the previous frame: its voxels may change again)
    launch_ray_keys(e->S, e->F, e->cur_depth, D, e->num_sms, e->stream);
    if (!(dbg & 1)) { rc = shard_barrier(e); if (rc != VH_OK) return rc; }
    launch_insert_keys(e->S, e->F, D, e->num_sms, e->stream);
  } else {
    if (!(dbg & 1)) { rc = shard_barrier(e); if (rc != VH_OK) return rc; }
    launch_alloc_visible(e->S, e->F, e->cur_depth, D, e->stream);              // replicated ray pass with an owner filter
  }
  if (s->rank == 0 && s->pull) CK(cudaEventRecord(s->released[b ^ 1], e->stream));   // every GPU has left the previous frame: its slot may be refilled
  CK(cudaEventRecord(e->ev[5], e->stream));
  launch_cull_list(e->S, e->F, D, e->num_sms, e->stream);
  CK(cudaEventRecord(e->ev[2], e->stream));
  e->S.weight_bound = ++e->integrate_launches + e->weight_bound_bias;
  pick_integrate_parts(e);
  launch_integrate(e->S, e->F, px, e->cur_rgb != nullptr, D, e->num_sms, e->stream);
  CK(cudaEventRecord(e->ev[3], e->stream));
  if (e->P.mc_per_frame) {
    if (!(dbg & 1)) { rc = shard_barrier(e); if (rc != VH_OK) return rc; }      // every GPU has integrated frame f
    DeviceView Dm = D;
    if (dbg & 2) Dm.peers = nullptr;
    launch_marching_cubes(e->S, e->F, Dm, D.visible, &D.counters->visible_count, 0, D.tri_offset, D.tri_count, e->num_sms, e->stream);
  }
  CK(cudaEventRecord(e->ev[4], e->stream));
  CK(cudaEventRecord(s->consumed[b], e->stream));
  s->used[b] = true;
  e->frames_in_flight++;
  return enqueue_readback(e);
}

int vh_integrate_sharded(vh_engine* e, const float* depth, const uint8_t* rgb, const float* c2w) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  return integrate_sharded_common(e, depth, rgb, c2w, true);
}
// the same with the frame already resident in rank 0's HBM (device pointers on rank 0; ignored elsewhere)
int vh_integrate_sharded_device(vh_engine* e, const float* d_depth, const uint8_t* d_rgb, const float* c2w) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  return integrate_sharded_common(e, d_depth, d_rgb, c2w, false);
}
// every GPU of the group is past everything enqueued before this call on every other GPU (collective, asynchronous): call it
// before anything outside the frame calls changes voxels that a peer's marching cubes may still be reading (vh_reset does).
int vh_shard_barrier(vh_engine* e) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  if (!e->shard) return fail(VH_ERR_INVALID, "vh_shard_connect has not been called");
  CK(cudaSetDevice(e->P.device));
  return shard_barrier(e);
}

// group-wide sums of the last frame's counters (collective)
int vh_shard_stats(vh_engine* e, vh_stats* sum) {
  if (!e || !sum) return fail(VH_ERR_INVALID, "null argument");
  vh_shard_state* s = e->shard;
  if (!s) return fail(VH_ERR_INVALID, "vh_shard_connect has not been called");
  vh_stats mine;
  int rc = vh_get_stats(e, &mine);
  if (rc != VH_OK) return rc;
  std::lock_guard<std::mutex> lk(e->mtx);
  unsigned long long h[6] = {mine.visible_blocks, mine.allocated_blocks, mine.voxel_updates, mine.triangles, mine.arena_triangles, 0};
  unsigned long long* d = nullptr;
  CK(cudaMalloc((void**)&d, sizeof(h)));
  CK(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, e->stream));
  NK(g_nccl.AllReduce(d, d, 6, ncclUint64, ncclSum, s->comm, e->stream));
  CK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(d);
  *sum = mine;
  sum->visible_blocks = (uint32_t)h[0]; sum->allocated_blocks = (uint32_t)h[1]; sum->voxel_updates = h[2]; sum->triangles = h[3];
  sum->arena_triangles = h[4];
  return VH_OK;
}

// Host-side merge of per-shard block lists into the reference's mesh order. part p holds nblocks[p] blocks with
// keys[p][3*i..] and counts[p][i] triangles, each part already in mesh order. out_part/out_index (length = total blocks)
// receive, for every position of the merged order, which part and which block of that part comes next.
int vh_mesh_order_merge(int n_parts, const int32_t* const* keys, const int* nblocks, int blocks_per_chunk, int32_t* out_part, int32_t* out_index) {
  if (n_parts <= 0 || !keys || !nblocks || !out_part || !out_index || blocks_per_chunk <= 0) return fail(VH_ERR_INVALID, "invalid argument");
  struct Head { int c[3]; int b[3]; };
  auto head_of = [&](int p, int i) {
    Head h;
    for (int a = 0; a < 3; a++) { h.b[a] = keys[p][3 * (size_t)i + a]; h.c[a] = (int)floorf((float)h.b[a] / (float)blocks_per_chunk); }   // block2chunk, tsdf.cu:256-260
    return h;
  };
  auto less = [](const Head& x, const Head& y) {
    for (int a = 0; a < 3; a++) if (x.c[a] != y.c[a]) return x.c[a] < y.c[a];
    for (int a = 0; a < 3; a++) if (x.b[a] != y.b[a]) return x.b[a] < y.b[a];
    return false;
  };
  std::vector<int> pos((size_t)n_parts, 0);
  size_t o = 0;
  for (;;) {
    int best = -1; Head bh{};
    for (int p = 0; p < n_parts; p++) {
      if (pos[p] >= nblocks[p]) continue;
      const Head h = head_of(p, pos[p]);
      if (best < 0 || less(h, bh)) { best = p; bh = h; }
    }
    if (best < 0) break;
    out_part[o] = best; out_index[o] = pos[best]; pos[best]++; o++;
  }
  return VH_OK;
}

// The whole map's mesh on rank 0, in the reference's order (collective). Other ranks get *n = 0.
int vh_shard_gather_mesh(vh_engine* e, int mode, vh_triangle* out, uint64_t cap, uint64_t* n_out) {
  if (!e) return fail(VH_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lk(e->mtx);
  vh_shard_state* s = e->shard;
  if (!s) return fail(VH_ERR_INVALID, "vh_shard_connect has not been called");
  CK(cudaSetDevice(e->P.device));
  CK(cudaStreamSynchronize(e->stream));
  int rc = finish_sync(e);
  if (rc != VH_OK) return rc;
  if (mode == VH_MESH_FULL_MAP) { rc = shard_barrier(e); if (rc != VH_OK) return rc; CK(cudaStreamSynchronize(e->stream)); }   // peers' voxels are final
  MeshBlocks mb;
  rc = collect_blocks(e, mode, mb);
  if (rc != VH_OK) { cudaFree(mb.tmp_arena); return rc; }
  const int nb = (int)mb.key.size();
  unsigned long long total = 0;
  for (int c : mb.cnt) total += (unsigned long long)c;
  std::vector<vh_triangle> tris((size_t)total);
  rc = gather_block_triangles(e, mb, tris.data(), total, nullptr);
  cudaFree(mb.tmp_arena);
  if (rc != VH_OK) return rc;
  if (mode == VH_MESH_FULL_MAP) { rc = shard_barrier(e); if (rc != VH_OK) return rc; }   // nobody integrates while a peer still meshes

  // sizes of every rank
  unsigned long long mine[2] = {(unsigned long long)nb, total};
  unsigned long long* d_sz = nullptr;
  CK(cudaMalloc((void**)&d_sz, sizeof(mine) * s->count));
  CK(cudaMemcpyAsync(d_sz + 2 * s->rank, mine, sizeof(mine), cudaMemcpyHostToDevice, e->stream));
  NK(g_nccl.AllGather(d_sz + 2 * s->rank, d_sz, 2, ncclUint64, s->comm, e->stream));
  std::vector<unsigned long long> sz((size_t)2 * s->count);
  CK(cudaMemcpyAsync(sz.data(), d_sz, sizeof(mine) * s->count, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  cudaFree(d_sz);

  // per-block records (x, y, z, count) and triangles travel as bytes
  std::vector<int32_t> rec((size_t)nb * 4);
  for (int i = 0; i < nb; i++) { int x, y, z; unpack_key(mb.key[i], x, y, z); rec[4 * (size_t)i] = x; rec[4 * (size_t)i + 1] = y; rec[4 * (size_t)i + 2] = z; rec[4 * (size_t)i + 3] = mb.cnt[i]; }
  if (s->rank != 0) {
    uint8_t* d_buf = nullptr;
    const size_t rb = rec.size() * sizeof(int32_t), tb = tris.size() * sizeof(vh_triangle);
    CK(cudaMalloc((void**)&d_buf, rb + tb + 16));
    if (rb) CK(cudaMemcpyAsync(d_buf, rec.data(), rb, cudaMemcpyHostToDevice, e->stream));
    if (tb) CK(cudaMemcpyAsync(d_buf + rb, tris.data(), tb, cudaMemcpyHostToDevice, e->stream));
    NK(g_nccl.Send(d_buf, rb + tb, ncclChar, 0, s->comm, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_buf);
    if (n_out) *n_out = 0;
    return VH_OK;
  }
  std::vector<std::vector<int32_t>> recs((size_t)s->count);
  std::vector<std::vector<vh_triangle>> parts((size_t)s->count);
  recs[0].swap(rec); parts[0].swap(tris);
  unsigned long long grand = sz[1];
  for (int q = 1; q < s->count; q++) {
    const size_t rb = (size_t)sz[2 * q] * 4 * sizeof(int32_t), tb = (size_t)sz[2 * q + 1] * sizeof(vh_triangle);
    grand += sz[2 * q + 1];
    uint8_t* d_buf = nullptr;
    CK(cudaMalloc((void**)&d_buf, rb + tb + 16));
    NK(g_nccl.Recv(d_buf, rb + tb, ncclChar, q, s->comm, e->stream));
    recs[q].resize((size_t)sz[2 * q] * 4); parts[q].resize((size_t)sz[2 * q + 1]);
    if (rb) CK(cudaMemcpyAsync(recs[q].data(), d_buf, rb, cudaMemcpyDeviceToHost, e->stream));
    if (tb) CK(cudaMemcpyAsync(parts[q].data(), d_buf + rb, tb, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(d_buf);
  }
  if (n_out) *n_out = grand;
  if (!out) return VH_OK;
  if (cap < grand) return fail(VH_ERR_INVALID, "output capacity %llu < %llu triangles", (unsigned long long)cap, grand);
  // merge by mesh order
  std::vector<std::vector<int32_t>> keys((size_t)s->count);
  std::vector<std::vector<unsigned long long>> first((size_t)s->count);
  std::vector<const int32_t*> kp((size_t)s->count);
  std::vector<int> nbs((size_t)s->count);
  size_t nblocks_all = 0;
  for (int q = 0; q < s->count; q++) {
    const size_t m = recs[q].size() / 4;
    keys[q].resize(m * 3); first[q].resize(m);
    unsigned long long acc = 0;
    for (size_t i = 0; i < m; i++) { for (int a = 0; a < 3; a++) keys[q][3 * i + a] = recs[q][4 * i + a]; first[q][i] = acc; acc += (unsigned long long)recs[q][4 * i + 3]; }
    kp[q] = keys[q].data(); nbs[q] = (int)m; nblocks_all += m;
  }
  std::vector<int32_t> op(nblocks_all), oi(nblocks_all);
  rc = vh_mesh_order_merge(s->count, kp.data(), nbs.data(), e->P.blocks_per_chunk, op.data(), oi.data());
  if (rc != VH_OK) return rc;
  size_t w = 0;
  for (size_t j = 0; j < nblocks_all; j++) {
    const int q = op[j]; const size_t i = (size_t)oi[j];
    const size_t cnt = (size_t)recs[q][4 * i + 3];
    if (cnt) memcpy(out + w, parts[q].data() + first[q][i], cnt * sizeof(vh_triangle));
    w += cnt;
  }
  return VH_OK;
}

}  // extern "C"
