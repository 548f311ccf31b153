// tests/emu/emu_engine.cpp — the whole per-frame path (pack -> allocate -> integrate -> marching cubes) with the product's
// own kernel sources executed on the CPU through tests/emu/cuda_runtime.h. TEST INFRASTRUCTURE ONLY (see emu_kernels.cpp):
// part of tests/emu/libvh_emu.so, loaded by tests/test_emu_engine.py; libvhsdf.so never links or loads it.
// The host side below restates the few lines of vh_create / reset_map / enqueue_stages that size and order the launches
// (voxel-hashing-sdf_b200/csrc/vh_engine.cu); the parameter blocks come from the engine's own vh_params_host.h.
#include <cuda_runtime.h>

#include "../../voxel-hashing-sdf_b200/csrc/vh_alloc.cu"
#include "../../voxel-hashing-sdf_b200/csrc/vh_mc.cu"
#include "../../voxel-hashing-sdf_b200/csrc/vh_stream.cu"
#include "../../voxel-hashing-sdf_b200/csrc/vh_params_host.h"

namespace vh {
// defined in emu_kernels.cpp (which includes vh_integrate.cu)
void emu_launch_pack(const float* depth, const uint8_t* rgb, uint2* out, int W, int H, float* tile_max, int* sched, FrameCounters* counters, uint32_t frame);
void emu_launch_integrate(const StaticParams& S, const FrameParams& F, const uint2* px, const DeviceView& D, bool color, int rev, int ctas);
}

using namespace vh;

struct emu_engine {
  vh_params P;
  StaticParams S;
  FrameParams F;
  DeviceView D;
  std::vector<u64> keys, key_heap;
  std::vector<int> slots, free_list, neg_count, sched, visible, tri_count;
  std::vector<uint32_t> stamps;
  std::vector<float> sdf, wgt, tile_max;
  std::vector<uchar4> rgb;
  std::vector<McWork> mc_queue;
  std::vector<uint4> work;
  std::vector<u64> inbox;
  int inbox_count[4] = {0, 0, 0, 0};
  int keys_done = 0;             // ray_keys_kernel of the current frame has run (emu_phase_keys)
  std::vector<vh_triangle> arena;
  std::vector<unsigned long long> tri_offset;
  std::vector<uint2> px;
  std::vector<uint4> tables;
  McQueueCtl mc_ctl[2];
  int free_top = 0, heap_counter = 0, map_error = 0, engine_error = 0, mc_parity = 0;
  uint32_t overflow_frame = 0;
  unsigned long long arena_top = 0, updates_total = 0;
  FrameCounters counters;
  uint64_t frames = 0;
  uint32_t integrate_launches = 0, weight_bound_bias = 0;
  int rev = 0, alloc_rev = 0;
  std::vector<int> full_list, full_cnt;          // full-map extraction (emu_full_map_mc)
  std::vector<unsigned long long> full_off;
  int full_valid = 0;
  PeerTable peers;               // multi-GPU emulation: every shard's tables and planes (plain host pointers here)
};

namespace {
struct AllocArgs { StaticParams S; FrameParams F; const float* depth; DeviceView D; int tiles_x; };
void run_alloc(void* p) { AllocArgs* a = static_cast<AllocArgs*>(p); alloc_visible_kernel(a->S, a->F, a->depth, a->D, a->tiles_x); }
struct KeysArgs { StaticParams S; FrameParams F; const float* depth; DeviceView D; int tiles_x, n_tiles; };
template <int TRX, int TRY> void run_ray_keys(void* p) { KeysArgs* a = static_cast<KeysArgs*>(p); ray_keys_kernel<TRX, TRY>(a->S, a->F, a->depth, a->D, a->tiles_x, a->n_tiles); }
struct InsertArgs { DeviceView D; uint32_t frame; };
void run_insert_keys(void* p) { InsertArgs* a = static_cast<InsertArgs*>(p); insert_keys_kernel(a->D, a->frame); }
struct McArgs { StaticParams S; uint32_t frame; DeviceView D; const int* list; const int* list_count; int full_map; unsigned long long* out_offset; int* out_count;
                McWork* queue; McQueueCtl* ctl; McQueueCtl* ctl_next; const uint4* tables; };
void run_filter_sharded(void* p) { McArgs* a = static_cast<McArgs*>(p); mc_filter_kernel<true>(a->S, a->frame, a->D, a->list, a->list_count, a->full_map, a->out_offset, a->out_count, a->queue, a->ctl); }
void run_mesh_sharded(void* p) { McArgs* a = static_cast<McArgs*>(p); if (a->S.mc_rev) return mc_mesh_kernel<true, true>(a->S, a->frame, a->D, a->full_map, a->out_offset, a->out_count, a->queue, a->ctl, a->ctl_next, a->tables); mc_mesh_kernel<true, false>(a->S, a->frame, a->D, a->full_map, a->out_offset, a->out_count, a->queue, a->ctl, a->ctl_next, a->tables); }
void run_filter(void* p) { McArgs* a = static_cast<McArgs*>(p); mc_filter_kernel<false>(a->S, a->frame, a->D, a->list, a->list_count, a->full_map, a->out_offset, a->out_count, a->queue, a->ctl); }
void run_mesh(void* p) { McArgs* a = static_cast<McArgs*>(p); if (a->S.mc_rev) return mc_mesh_kernel<false, true>(a->S, a->frame, a->D, a->full_map, a->out_offset, a->out_count, a->queue, a->ctl, a->ctl_next, a->tables); mc_mesh_kernel<false, false>(a->S, a->frame, a->D, a->full_map, a->out_offset, a->out_count, a->queue, a->ctl, a->ctl_next, a->tables); }
// test-only kernel: the step-by-step DDA (ray_march) and the merge formulation (merge_fill_keys) of the same tile of rays,
// compared key by key; out[0] += mismatching steps, out[1] += steps compared, out[2] += non-empty keys
struct MarchCmpArgs { StaticParams S; FrameParams F; const float* depth; int tiles_x; unsigned long long* out; };
void run_march_compare(void* p) {
  MarchCmpArgs* a = static_cast<MarchCmpArgs*>(p);
  const StaticParams& S = a->S;
  const int K = S.max_steps, tid = threadIdx.x;
  u64* skeys = reinterpret_cast<u64*>(emu::g_cta->dyn_smem);
  float* sT = reinterpret_cast<float*>(skeys + (size_t)K * RAYS);
  u64* seq = reinterpret_cast<u64*>(sT + (size_t)RAYS * 3 * K);      // [K][RAYS] keys of the sequential march
  static int s_death[RAYS];
  const int tile_x = blockIdx.x % a->tiles_x, tile_y = blockIdx.x / a->tiles_x;
  merge_fill_keys_t<RAYS_X, RAYS_Y>(S, a->F, a->depth, tile_x, tile_y, skeys, sT, s_death);
  if (tid < RAYS) {
    RayState R;
    ray_setup(S, a->F, a->depth, tile_x * RAYS_X + (tid & (RAYS_X - 1)), tile_y * RAYS_Y + (tid / RAYS_X), R);
    ray_march(R, K, seq + tid);
    for (int s = 0; s < K; s++) {
      const u64 m = s <= s_death[tid] ? skeys[s * RAYS + tid] : KEY_EMPTY;
      a->out[0] += m != seq[s * RAYS + tid];
      a->out[1] += 1;
      a->out[2] += seq[s * RAYS + tid] != KEY_EMPTY;
    }
  }
}
}  // namespace

extern "C" {

// the two DDA formulations over every ray of one frame: out3 = {mismatching steps, steps compared, non-empty keys}
void emu_compare_march(const vh_params* p, const float* depth, const float* c2w, unsigned long long* out3) {
  StaticParams S; memset(&S, 0, sizeof(S));
  derive_static_params(*p, S);
  FrameParams F; memset(&F, 0, sizeof(F));
  derive_frame_params(*p, S, c2w, F);
  F.frame = 1;
  out3[0] = out3[1] = out3[2] = 0;
  const int tiles_x = (S.nrx + RAYS_X - 1) / RAYS_X, tiles_y = (S.nry + RAYS_Y - 1) / RAYS_Y;
  MarchCmpArgs a{S, F, depth, tiles_x, out3};
  const size_t smem = (size_t)S.max_steps * RAYS * sizeof(u64) * 2 + (size_t)RAYS * 3 * S.max_steps * sizeof(float);
  emu::run_grid(dim3(tiles_x * tiles_y), dim3(ALLOC_THREADS), run_march_compare, &a, smem);
}

emu_engine* emu_create(const vh_params* p, int integrate_rev, int cull, int exact_color, int alloc_rev, int mc_rev) {
  if (!p || p->voxels_per_block != VPB || p->shard_count < 1 || p->shard_count > MAX_SHARDS) return nullptr;
  emu_engine* e = new emu_engine;
  e->P = *p;
  StaticParams& S = e->S; memset(&S, 0, sizeof(S));
  derive_static_params(*p, S);
  S.verify = 0; S.integrate_ctas_per_sm = 4; S.integrate_cull = cull; S.integrate_parts = 2; S.integrate_rev = integrate_rev;
  S.mc_rev = mc_rev;
  e->rev = integrate_rev; e->alloc_rev = alloc_rev; e->weight_bound_bias = exact_color ? 1u << 20 : 0u;
  uint64_t want = (uint64_t)p->num_buckets * (uint64_t)p->entries_per_bucket, cap = 1024;      // vh_create
  while (cap < want) cap <<= 1;
  const size_t nb = (size_t)p->pool_blocks, rays = (size_t)S.nrx * S.nry;
  DeviceView& D = e->D; memset(&D, 0, sizeof(D));
  D.list_cap = (int)std::max(rays * (size_t)S.max_steps, nb);
  e->keys.assign(cap, KEY_EMPTY); e->slots.assign(cap, -1); e->stamps.assign(cap, 0u);          // reset_map
  e->free_list.resize(nb); for (size_t i = 0; i < nb; i++) e->free_list[i] = (int)(nb - 1 - i);   // init_free_list_kernel
  e->free_top = (int)nb; e->key_heap.assign(nb, 0);
  e->sdf.assign(nb * BLOCK_VOX, 0.0f); e->wgt.assign(nb * BLOCK_VOX, 0.0f);
  if (S.use_color) e->rgb.assign(nb * BLOCK_VOX, uchar4{0, 0, 0, 0});
  e->neg_count.assign(nb, 0); e->sched.assign(9 * 32, 0);
  e->tile_max.assign((size_t)((p->width + 15) / 16) * ((p->height + 15) / 16), 0.0f);
  e->visible.assign(D.list_cap, 0); e->mc_queue.resize(D.list_cap); e->work.resize(D.list_cap);
  D.inbox_cap = (int)std::max<size_t>(rays * (size_t)S.max_steps, 1024);
  e->inbox.assign(2 * (size_t)D.inbox_cap, 0);
  e->arena.resize(std::max<size_t>((size_t)p->tri_arena_bytes / sizeof(vh_triangle), 1024));
  e->tri_offset.assign(nb, 0); e->tri_count.assign(nb, 0);
  e->px.assign((size_t)p->width * p->height + 1, make_uint2(0u, 0u));                          // + the sentinel record
  memset(e->mc_ctl, 0, sizeof(e->mc_ctl)); memset(&e->counters, 0, sizeof(e->counters));
  {   // upload_mc_tables: expanded case tables as the mesh kernel copies them (tri 4096 B, then ntri 256 B)
    static signed char tri[256][16]; static unsigned char ntri[256]; static unsigned short em[256];
    vh_mc_expand_tables(tri, ntri, em);
    e->tables.resize((sizeof(tri) + sizeof(ntri)) / sizeof(uint4));
    memcpy(e->tables.data(), tri, sizeof(tri)); memcpy(reinterpret_cast<char*>(e->tables.data()) + sizeof(tri), ntri, sizeof(ntri));
  }
  D.map.keys = e->keys.data(); D.map.slots = e->slots.data(); D.map.mask = (uint32_t)cap - 1; D.map.free_list = e->free_list.data();
  D.map.free_top = &e->free_top; D.map.key_heap = e->key_heap.data(); D.map.heap_counter = &e->heap_counter; D.map.error_flag = &e->map_error;
  D.map.num_blocks = p->pool_blocks;
  D.stamps = e->stamps.data(); D.sdf = e->sdf.data(); D.wgt = e->wgt.data(); D.rgb = S.use_color ? e->rgb.data() : nullptr;
  D.sched = e->sched.data(); D.work = e->work.data(); D.inbox = e->inbox.data(); D.inbox_count = e->inbox_count; D.inbox_done = e->inbox_count + 2; D.tile_max = e->tile_max.data(); D.neg_count = e->neg_count.data(); D.visible = e->visible.data();
  D.counters = &e->counters; D.arena = e->arena.data(); D.arena_top = &e->arena_top; D.arena_cap = e->arena.size();
  D.tri_offset = e->tri_offset.data(); D.tri_count = e->tri_count.data(); D.engine_error = &e->engine_error; D.overflow_frame = &e->overflow_frame;
  D.updates_total = &e->updates_total; D.peers = nullptr; D.mc_queue = e->mc_queue.data(); D.mc_ctl = e->mc_ctl; D.mc_parity = &e->mc_parity;
  return e;
}

void emu_destroy(emu_engine* e) { delete e; }

// One frame in two phases. Single map: the order of enqueue_stages for frames resident on the device — pack (resets the
// counters), allocate, integrate; then marching cubes over the visible list. Sharded map (vh_integrate_sharded): every
// rank runs phase 1 on the broadcast frame, a barrier, every rank runs phase 2 reading its peers' tables and planes.
// allocation revision 2, first half (launch_ray_keys): the frame's keys into their owners' inboxes. On a sharded map every rank
// runs this before any rank inserts (the frame barrier of vh_integrate_sharded); tile shapes alternate between frames.
static void emu_begin_frame(emu_engine* e, const float* c2w) {
  derive_frame_params(e->P, e->S, c2w, e->F);
  e->F.frame = (uint32_t)(++e->frames);
}
static void emu_ray_keys(emu_engine* e, const float* depth) {
  const StaticParams& S = e->S;
  const int shape = (int)(e->frames % 3);
  const int trx = shape == 0 ? 4 : 2, try_ = shape == 2 ? 1 : 2;
  const int tiles_x = (S.nrx + trx - 1) / trx, tiles_y = (S.nry + try_ - 1) / try_, n_tiles = tiles_x * tiles_y;
  const bool routed = S.shard_count > 1 && e->D.peers != nullptr;
  const int grid = routed ? (n_tiles + (int)S.shard_count - 1) / (int)S.shard_count : n_tiles;
  KeysArgs a{S, e->F, depth, e->D, tiles_x, n_tiles};
  const size_t smem = ray_keys_smem_bytes(S.max_steps, trx * try_);
  emu::run_grid(dim3(grid), dim3(KEYS_THREADS), shape == 0 ? run_ray_keys<4, 2> : shape == 1 ? run_ray_keys<2, 2> : run_ray_keys<2, 1>, &a, smem);
}
int emu_phase_keys(emu_engine* e, const float* depth, const float* c2w) {
  if (e->alloc_rev != 2) return 0;
  emu_begin_frame(e, c2w);
  emu_ray_keys(e, depth);
  e->keys_done = 1;
  return e->map_error;
}

int emu_phase_integrate(emu_engine* e, const float* depth, const uint8_t* rgb, const float* c2w) {
  const StaticParams& S = e->S; DeviceView& D = e->D;
  if (!e->keys_done) emu_begin_frame(e, c2w);
  const uint8_t* rgb_in = S.use_color ? rgb : nullptr;
  emu_launch_pack(depth, rgb_in, e->px.data(), S.W, S.H, D.tile_max, D.sched, D.counters, e->F.frame);
  if (e->alloc_rev == 2) {
    if (!e->keys_done) emu_ray_keys(e, depth);
    e->keys_done = 0;
    InsertArgs ia{D, e->F.frame};
    emu::run_grid(dim3(3), dim3(INSERT_THREADS), run_insert_keys, &ia);                          // launch_insert_keys
  } else {
    const int tiles_x = (S.nrx + RAYS_X - 1) / RAYS_X, tiles_y = (S.nry + RAYS_Y - 1) / RAYS_Y;     // launch_alloc_visible
    const size_t smem = 2 * CHUNK_KEYS * sizeof(u64) + (size_t)S.max_steps * RAYS * sizeof(int);
    AllocArgs a{S, e->F, depth, D, tiles_x};
    emu::run_grid(dim3(tiles_x * tiles_y), dim3(ALLOC_THREADS), run_alloc, &a, smem);
  }
  e->S.weight_bound = ++e->integrate_launches + e->weight_bound_bias;
  emu_launch_integrate(e->S, e->F, e->px.data(), D, rgb_in != nullptr, e->rev, 2);
  return e->map_error | (e->engine_error << 8);
}

int emu_phase_mc(emu_engine* e) {
  const StaticParams& S = e->S; DeviceView& D = e->D;
  if (e->P.mc_per_frame) {
    McQueueCtl* ctl = D.mc_ctl + (e->mc_parity & 1);                                                // launch_marching_cubes
    McQueueCtl* ctl_next = D.mc_ctl + ((e->mc_parity & 1) ^ 1);
    e->mc_parity ^= 1;
    McArgs m{S, e->F.frame, D, D.visible, &D.counters->visible_count, 0, D.tri_offset, D.tri_count, D.mc_queue, ctl, ctl_next, e->tables.data()};
    const bool sharded = S.shard_count > 1 && D.peers;
    emu::run_grid(dim3(3), dim3(256), sharded ? run_filter_sharded : run_filter, &m);
    emu::run_grid(dim3(3), dim3(MC_THREADS), sharded ? run_mesh_sharded : run_mesh, &m, (size_t)MC_WARPS * TILE_PAD * sizeof(uint32_t));
  }
  return e->map_error | (e->engine_error << 8);
}

int emu_process_frame(emu_engine* e, const float* depth, const uint8_t* rgb, const float* c2w) {
  const int rc = emu_phase_integrate(e, depth, rgb, c2w);
  return rc | emu_phase_mc(e);
}

// vh_extract_mesh(VH_MESH_FULL_MAP): list every allocated block (list_all_blocks_kernel), then the two marching-cubes kernels with
// full_map = 1 into separate offset/count arrays (the per-frame meshes stay). Returns the triangle count.
struct ListArgs { DeviceView D; int* list; int* list_count; };
static void run_list_all(void* p) { ListArgs* a = static_cast<ListArgs*>(p); list_all_blocks_kernel(a->D, a->list, a->list_count); }

long long emu_full_map_mc(emu_engine* e) {
  const StaticParams& S = e->S; DeviceView& D = e->D;
  const int nb = D.map.num_blocks;
  e->full_list.assign(nb, 0); e->full_off.assign(nb, 0); e->full_cnt.assign(nb, 0);
  int count = 0;
  ListArgs la{D, e->full_list.data(), &count};
  emu::run_grid(dim3((nb + 255) / 256), dim3(256), run_list_all, &la);
  McQueueCtl* ctl = D.mc_ctl + (e->mc_parity & 1);
  McQueueCtl* ctl_next = D.mc_ctl + ((e->mc_parity & 1) ^ 1);
  e->mc_parity ^= 1;
  McArgs m{S, e->F.frame, D, e->full_list.data(), &count, 1, e->full_off.data(), e->full_cnt.data(), D.mc_queue, ctl, ctl_next, e->tables.data()};
  const bool sharded = S.shard_count > 1 && D.peers;
  emu::run_grid(dim3(3), dim3(256), sharded ? run_filter_sharded : run_filter, &m);
  emu::run_grid(dim3(3), dim3(MC_THREADS), sharded ? run_mesh_sharded : run_mesh, &m, (size_t)MC_WARPS * TILE_PAD * sizeof(uint32_t));
  long long t = 0;
  for (int i = 0; i < nb; i++) t += e->full_cnt[i];
  e->full_valid = 1;
  return (e->engine_error & 1) ? -1 : t;
}

// ---- out-of-core tier (csrc/vh_stream.cu): the kernels, with the few host lines of vh_far_blocks / vh_evict_blocks / vh_upload_blocks restated
struct FarArgs { StaticParams S; FrameParams F; DeviceView D; u64* out; int cap; int* count; };
static void run_far(void* p) { FarArgs* a = static_cast<FarArgs*>(p); far_blocks_kernel(a->S, a->F, a->D, a->out, a->cap, a->count); }
struct EvictArgs { DeviceView D; const u64* keys; int n; int* released; };
static void run_evict(void* p) { EvictArgs* a = static_cast<EvictArgs*>(p); evict_kernel(a->D, a->keys, a->n, a->released); }
struct UpArgs { DeviceView D; const u64* keys; int n; int* slots; const float* sdf; const float* wgt; const uint8_t* rgb; };
static void run_up_insert(void* p) { UpArgs* a = static_cast<UpArgs*>(p); upload_insert_kernel(a->D, a->keys, a->n, a->slots); }
static void run_up_scatter(void* p) { UpArgs* a = static_cast<UpArgs*>(p); upload_scatter_kernel(a->D, a->slots, a->n, a->sdf, a->wgt, a->rgb); }
struct RebuildArgs { DeviceView D; uint32_t capacity; u64* k; int* s; uint32_t* st; int room; int* count; int n; };
static void run_rebuild_collect(void* p) { RebuildArgs* a = static_cast<RebuildArgs*>(p); rebuild_collect_kernel(a->D, a->capacity, a->k, a->s, a->st, a->room, a->count); }
static void run_rebuild_insert(void* p) { RebuildArgs* a = static_cast<RebuildArgs*>(p); rebuild_insert_kernel(a->D, a->k, a->s, a->st, a->n); }

int emu_far_blocks(emu_engine* e, const float* c2w, int* out_xyz, int cap) {
  FrameParams F; derive_frame_params(e->P, e->S, c2w, F); F.frame = 0;
  std::vector<u64> out(std::max(cap, 1));
  int count = 0;
  FarArgs a{e->S, F, e->D, out.data(), out_xyz ? cap : 0, &count};
  emu::run_grid(dim3((e->D.map.num_blocks + 255) / 256), dim3(256), run_far, &a);
  const int m = std::min(count, out_xyz ? cap : 0);
  std::sort(out.begin(), out.begin() + m);
  for (int i = 0; i < m; i++) unpack_key(out[i], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
  return count;
}

int emu_rebuild_table(emu_engine* e) {
  const uint32_t capacity = e->D.map.mask + 1;
  const int room = (int)std::min<size_t>(capacity, 2 * (size_t)e->P.pool_blocks);
  std::vector<u64> k(room); std::vector<int> s(room); std::vector<uint32_t> st(room);
  int count = 0;
  RebuildArgs a{e->D, capacity, k.data(), s.data(), st.data(), room, &count, 0};
  emu::run_grid(dim3((capacity + 255) / 256), dim3(256), run_rebuild_collect, &a);
  if (count > room) return -1;
  std::fill(e->keys.begin(), e->keys.end(), KEY_EMPTY); std::fill(e->slots.begin(), e->slots.end(), -1); std::fill(e->stamps.begin(), e->stamps.end(), 0u);
  a.n = count;
  if (count > 0) emu::run_grid(dim3((count + 255) / 256), dim3(256), run_rebuild_insert, &a);
  e->counters.visible_count = 0;
  return count;
}

void emu_get_blocks(const emu_engine* e, const int* keys_xyz, int n, float* sdf, float* w, uint8_t* rgb, uint8_t* found, int* neg);
// returns the number of blocks released; found[i] / planes as vh_download_blocks; the last frame's visible list is void afterwards
int emu_evict_blocks(emu_engine* e, const int* keys_xyz, int n, float* sdf, float* w, uint8_t* rgb, uint8_t* found) {
  std::vector<int> neg(std::max(n, 1));
  emu_get_blocks(e, keys_xyz, n, sdf, w, rgb, found, neg.data());
  std::vector<u64> packed(std::max(n, 1));
  for (int i = 0; i < n; i++) packed[i] = pack_key(keys_xyz[3 * i], keys_xyz[3 * i + 1], keys_xyz[3 * i + 2]);
  int released = 0;
  EvictArgs a{e->D, packed.data(), n, &released};
  emu::run_grid(dim3((n + 7) / 8), dim3(256), run_evict, &a);
  if (released > 0) {      // key_heap: drop the released keys, keep the insertion order (host side, as vh_evict_blocks does)
    const std::unordered_set<u64> gone(packed.begin(), packed.begin() + n);
    int m = 0;
    for (int i = 0; i < e->heap_counter; i++) if (!gone.count(e->key_heap[i])) e->key_heap[m++] = e->key_heap[i];
    if (m != e->heap_counter - released) return -1;
    e->heap_counter = m;
  }
  e->counters.visible_count = 0;
  return released;
}

int emu_upload_blocks(emu_engine* e, const int* keys_xyz, int n, const float* sdf, const float* w, const uint8_t* rgb) {
  std::vector<u64> packed(std::max(n, 1));
  for (int i = 0; i < n; i++) packed[i] = pack_key(keys_xyz[3 * i], keys_xyz[3 * i + 1], keys_xyz[3 * i + 2]);
  std::vector<int> slots(std::max(n, 1), -2);
  UpArgs a{e->D, packed.data(), n, slots.data(), sdf, w, rgb};
  emu::run_grid(dim3((n + 255) / 256), dim3(256), run_up_insert, &a);
  emu::run_grid(dim3((n + 7) / 8), dim3(256), run_up_scatter, &a);
  return e->map_error;
}

int emu_free_slots(const emu_engine* e) { return e->free_top; }

// vh_shard_connect: every rank gets a view of every rank's table and planes (CUDA-IPC mappings on the GPU, pointers here)
int emu_connect(emu_engine** ranks, int n) {
  if (n < 1 || n > MAX_SHARDS) return -1;
  for (int r = 0; r < n; r++) if (!ranks[r] || (int)ranks[r]->S.shard_count != n || (int)ranks[r]->S.shard_rank != r) return -2;
  for (int r = 0; r < n; r++) {
    PeerTable& T = ranks[r]->peers; memset(&T, 0, sizeof(T));
    for (int q = 0; q < n; q++) {
      const DeviceView& Q = ranks[q]->D;
      T.v[q].keys = Q.map.keys; T.v[q].slots = Q.map.slots; T.v[q].stamps = Q.stamps; T.v[q].neg_count = Q.neg_count; T.v[q].sdf = Q.sdf; T.v[q].rgb = Q.rgb;
      T.v[q].mask = Q.map.mask; T.v[q].inbox = Q.inbox; T.v[q].inbox_count = Q.inbox_count;
    }
    ranks[r]->D.peers = &ranks[r]->peers;
  }
  return 0;
}

int emu_num_visible(const emu_engine* e) { return e->counters.visible_count; }
unsigned long long emu_last_updates(const emu_engine* e) { return e->counters.voxel_updates; }
unsigned long long emu_last_triangles(const emu_engine* e) { return e->counters.triangles; }
int emu_num_blocks(const emu_engine* e) { return e->heap_counter; }

void emu_visible_keys(const emu_engine* e, int* out_xyz) {
  for (int i = 0; i < e->counters.visible_count; i++) unpack_key(e->keys[e->visible[i]], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
}
void emu_all_keys(const emu_engine* e, int* out_xyz) {
  for (int i = 0; i < e->heap_counter; i++) unpack_key(e->key_heap[i], out_xyz[3 * i], out_xyz[3 * i + 1], out_xyz[3 * i + 2]);
}
static int find_slot(const emu_engine* e, int x, int y, int z) {
  const u64 key = pack_key(x, y, z);
  uint32_t h = hash_key(key) & e->D.map.mask;
  for (;;) { const u64 cur = e->keys[h]; if (cur == key) return e->slots[h]; if (cur == KEY_EMPTY) return -1; h = (h + 1) & e->D.map.mask; }
}
// voxels of the given blocks: sdf, weight [n][512], rgb [n][512][3], found [n]; neg [n] = the kernel-maintained negative-voxel counter
void emu_get_blocks(const emu_engine* e, const int* keys_xyz, int n, float* sdf, float* w, uint8_t* rgb, uint8_t* found, int* neg) {
  for (int i = 0; i < n; i++) {
    const int s = find_slot(e, keys_xyz[3 * i], keys_xyz[3 * i + 1], keys_xyz[3 * i + 2]);
    found[i] = s >= 0;
    if (s < 0) continue;
    memcpy(sdf + (size_t)i * 512, &e->sdf[(size_t)s * 512], 512 * sizeof(float));
    memcpy(w + (size_t)i * 512, &e->wgt[(size_t)s * 512], 512 * sizeof(float));
    neg[i] = e->neg_count[s];
    if (e->S.use_color) for (int v = 0; v < 512; v++) { const uchar4 c = e->rgb[(size_t)s * 512 + v]; uint8_t* o = rgb + ((size_t)i * 512 + v) * 3; o[0] = c.x; o[1] = c.y; o[2] = c.z; }
  }
}
// stored triangles of the given blocks, concatenated in the given order: returns the count; xyz [T][3][3], rgb [T][3][3] (may be null to count)
long long emu_block_triangles(const emu_engine* e, const int* keys_xyz, int n, float* xyz, uint8_t* rgb, int full_map) {
  long long t = 0;
  if (full_map && !e->full_valid) return -1;
  for (int i = 0; i < n; i++) {
    const int s = find_slot(e, keys_xyz[3 * i], keys_xyz[3 * i + 1], keys_xyz[3 * i + 2]);
    if (s < 0) continue;
    const int cnt = full_map ? e->full_cnt[s] : e->tri_count[s];
    const unsigned long long off = full_map ? e->full_off[s] : e->tri_offset[s];
    for (int k = 0; k < cnt; k++, t++) {
      if (!xyz) continue;
      const vh_triangle& T = e->arena[off + k];
      for (int v = 0; v < 3; v++) {
        xyz[(t * 3 + v) * 3 + 0] = T.p[v].x; xyz[(t * 3 + v) * 3 + 1] = T.p[v].y; xyz[(t * 3 + v) * 3 + 2] = T.p[v].z;
        rgb[(t * 3 + v) * 3 + 0] = T.p[v].r; rgb[(t * 3 + v) * 3 + 1] = T.p[v].g; rgb[(t * 3 + v) * 3 + 2] = T.p[v].b;
      }
    }
  }
  return t;
}

}  // extern "C"
