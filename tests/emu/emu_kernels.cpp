// tests/emu/emu_kernels.cpp — runs the engine's integrate/pack kernels (the product's own source, included as text) on
// the CPU through tests/emu/cuda_runtime.h. TEST INFRASTRUCTURE ONLY: built by tests/emu/Makefile into
// tests/emu/libvh_emu.so and loaded by tests/test_emu_integrate.py; libvhsdf.so never links or loads it.
#include <cuda_runtime.h>
#include "emu_runtime.h"

#include "../../voxel-hashing-sdf_b200/csrc/vh_integrate.cu"

namespace {

struct IntegrateArgs { vh::StaticParams S; vh::FrameParams F; const uint2* px; vh::DeviceView D; int variant; };

template <bool C, bool V, bool DL, int M>
void run_direct(void* p) {
  IntegrateArgs* a = static_cast<IntegrateArgs*>(p);
  vh::integrate_kernel_direct<C, V, DL, M, false>(a->S, a->F, a->px, a->D);
}
template <bool C, bool V, bool DL>
void run_direct_wide(void* p) {      // 64-bit voxel indices (pools beyond 2^23 blocks)
  IntegrateArgs* a = static_cast<IntegrateArgs*>(p);
  vh::integrate_kernel_direct<C, V, DL, 3, true>(a->S, a->F, a->px, a->D);
}
template <bool C, bool V, bool DL, int PARTS>
void run_staged(void* p) {
  IntegrateArgs* a = static_cast<IntegrateArgs*>(p);
  vh::integrate_kernel_staged<C, V, DL, PARTS, 4>(a->S, a->F, a->px, a->D);
}
void run_cull_list(void* p) {
  IntegrateArgs* a = static_cast<IntegrateArgs*>(p);
  vh::cull_list_kernel(a->S, a->F, a->D);
}

struct PackArgs { const float* depth; const uint8_t* rgb; uint2* out; int W, H; float* tile_max; int* sched; vh::FrameCounters* counters; uint32_t frame; };
void run_pack(void* p) {
  PackArgs* a = static_cast<PackArgs*>(p);
  vh::pack_frame_kernel(a->depth, a->rgb, a->out, a->W, a->H, a->tile_max, a->sched, a->counters, a->frame, 0);
}

}  // namespace

namespace vh {

void emu_launch_pack(const float* depth, const uint8_t* rgb, uint2* out, int W, int H, float* tile_max, int* sched, FrameCounters* counters, uint32_t frame) {
  PackArgs pa{depth, rgb, out, W, H, tile_max, sched, counters, frame};
  emu::run_grid(dim3((W + 15) / 16, (H + 15) / 16), dim3(16, 16), run_pack, &pa);
}

// launch_cull_list + launch_integrate (csrc/vh_integrate.cu) on `ctas` emulated CTAs; rev 2 = staged (two steps at a time), else direct
void emu_launch_integrate(const StaticParams& S, const FrameParams& F, const uint2* px, const DeviceView& D, bool color, int rev, int ctas) {
  color = color && S.use_color;
  IntegrateArgs ia{S, F, px, D, rev};
  const bool delta = S.weight_bound <= 65536u;
  emu::run_grid(dim3(2), dim3(256), run_cull_list, &ia);
  if (rev == 2) {
    // whole-block and half-block work items alternate between frames
    const bool halves = (F.frame & 1u) != 0;
    void (*entry)(void*) = halves ? (!color ? run_staged<false, false, false, 2> : delta ? run_staged<true, false, true, 2> : run_staged<true, false, false, 2>)
                                  : (!color ? run_staged<false, false, false, 1> : delta ? run_staged<true, false, true, 1> : run_staged<true, false, false, 1>);
    emu::run_grid(dim3(std::max(ctas, 1)), dim3(STG_THREADS), entry, &ia, integrate_staged_smem_bytes(halves ? 2 : 1));
  } else {
    void (*entry)(void*) = !color ? run_direct<false, false, false, 3> : delta ? run_direct<true, false, true, 3> : run_direct<true, false, false, 3>;
    emu::run_grid(dim3(std::max(ctas, 1)), dim3(INT_THREADS), entry, &ia);
  }
}

}  // namespace vh

extern "C" {

struct emu_integrate_io {
  int W, H;
  float fx, fy, cx, cy, max_depth, vox_size, trunc;
  int use_color, cull, two_steps, verify, exact_color, ctas, variant;   // ctas: emulated grid size in CTAs; variant: 1 = direct kernel, 2 = staged, 3 = direct with 64-bit voxel indices
  unsigned weight_bound, frame, rcp_seed;
  const float* c2w;              // [16]
  const float* depth;            // [H*W]
  const uint8_t* rgb;            // [H*W*3] or null
  int n_visible;
  const int* keys_xyz;           // [n][3] block coordinates of the frame's visible list
  const int* slots;              // [n] pool slot of each
  float* sdf; float* wgt; uint8_t* rgb4; int* neg_count;   // planes: [pool*512] f32, f32, u8x4; [pool]
  unsigned long long voxel_updates, culled, mismatch, collectives, slow_steps;   // out
  int engine_error;              // out
};

int emu_integrate(emu_integrate_io* io) {
  using namespace vh;
  emu::g_rcp_seed = io->rcp_seed;
  StaticParams S; memset(&S, 0, sizeof(S));
  S.W = io->W; S.H = io->H; S.fx = io->fx; S.fy = io->fy; S.cx = io->cx; S.cy = io->cy; S.max_depth = io->max_depth;
  S.vox_size = io->vox_size; S.trunc = io->trunc; S.use_color = io->use_color;
  S.round_eps = 7.5e-7f * (float)std::max(io->W, io->H) + 2e-5f;    // vh_create, csrc/vh_engine.cu
  S.byte_bias = 0x4B000000u; S.verify = io->verify; S.weight_bound = io->weight_bound; S.integrate_cull = io->cull; S.integrate_parts = io->two_steps ? 1 : 2;
  FrameParams F; memset(&F, 0, sizeof(F));
  memcpy(F.c2w, io->c2w, sizeof(F.c2w)); F.frame = io->frame;

  const int n = io->n_visible;
  std::vector<u64> keys(std::max(n, 1));
  std::vector<int> visible(std::max(n, 1));
  for (int i = 0; i < n; i++) { keys[i] = pack_key(io->keys_xyz[3 * i], io->keys_xyz[3 * i + 1], io->keys_xyz[3 * i + 2]); visible[i] = i; }
  const int tiles = ((io->W + 15) / 16) * ((io->H + 15) / 16);
  std::vector<float> tile_max(tiles, -1.0f);
  std::vector<int> sched((NSCHED + 1) * 32, 12345);    // pack_frame_kernel must zero the counters it owns
  std::vector<uint4> work(std::max(n, 1));
  std::vector<uint2> px((size_t)io->W * io->H + 1, make_uint2(0u, 0u));   // + the sentinel record (vh_create)
  FrameCounters counters; memset(&counters, 0xAB, sizeof(counters));
  int engine_error = 0; unsigned long long updates_total = 0;

  PackArgs pa{io->depth, io->use_color ? io->rgb : nullptr, px.data(), io->W, io->H, tile_max.data(), sched.data(), &counters, io->frame};
  emu::run_grid(dim3((io->W + 15) / 16, (io->H + 15) / 16), dim3(16, 16), run_pack, &pa);
  if (counters.visible_count != 0 || counters.frame != io->frame || counters.voxel_updates != 0) return -2;
  counters.visible_count = n;     // the allocation pass would have counted the list

  DeviceView D; memset(&D, 0, sizeof(D));
  D.map.keys = keys.data(); D.map.slots = const_cast<int*>(io->slots);
  D.sdf = io->sdf; D.wgt = io->wgt; D.rgb = reinterpret_cast<uchar4*>(io->rgb4); D.neg_count = io->neg_count;
  D.sched = sched.data(); D.work = work.data(); D.tile_max = tile_max.data(); D.visible = visible.data(); D.list_cap = std::max(n, 1);
  D.counters = &counters; D.engine_error = &engine_error; D.updates_total = &updates_total;

  IntegrateArgs ia{S, F, px.data(), D, 0};
  const bool color = io->use_color != 0, delta = !io->exact_color && S.weight_bound <= 65536u;
  void (*entry)(void*) = nullptr;
  emu::g_collectives = 0;
  emu::run_grid(dim3(2), dim3(256), run_cull_list, &ia);
  if (io->variant == 2) {     // integrate_kernel_staged; two_steps = 1 selects whole-block work items, 0 = x-halves
#define PICK2(C, V, DL) (io->two_steps ? run_staged<C, V, DL, 1> : run_staged<C, V, DL, 2>)
    if (io->verify) entry = !color ? PICK2(false, true, false) : delta ? PICK2(true, true, true) : PICK2(true, true, false);
    else entry = !color ? PICK2(false, false, false) : delta ? PICK2(true, false, true) : PICK2(true, false, false);
#undef PICK2
    emu::run_grid(dim3(std::max(io->ctas, 1)), dim3(STG_THREADS), entry, &ia, integrate_staged_smem_bytes(io->two_steps ? 1 : 2));
  } else if (io->variant == 3) {      // integrate_kernel_direct with 64-bit voxel indices
    if (io->verify) entry = !color ? run_direct_wide<false, true, false> : delta ? run_direct_wide<true, true, true> : run_direct_wide<true, true, false>;
    else entry = !color ? run_direct_wide<false, false, false> : delta ? run_direct_wide<true, false, true> : run_direct_wide<true, false, false>;
    emu::run_grid(dim3(std::max(io->ctas, 1)), dim3(INT_THREADS), entry, &ia);
  } else {                    // integrate_kernel_direct; two_steps selects the 4-CTA (64-register) build
#define PICK1(C, V, DL) (io->two_steps ? run_direct<C, V, DL, 4> : run_direct<C, V, DL, 3>)
    if (io->verify) entry = !color ? PICK1(false, true, false) : delta ? PICK1(true, true, true) : PICK1(true, true, false);
    else entry = !color ? PICK1(false, false, false) : delta ? PICK1(true, false, true) : PICK1(true, false, false);
#undef PICK1
    emu::run_grid(dim3(std::max(io->ctas, 1)), dim3(INT_THREADS), entry, &ia);
  }
  io->voxel_updates = counters.voxel_updates; io->culled = counters.pad[1]; io->mismatch = counters.pad[0];
  io->collectives = emu::g_collectives; io->slow_steps = counters.pad[2]; io->engine_error = engine_error;
  return updates_total == counters.voxel_updates ? 0 : -3;
}

}  // extern "C"
