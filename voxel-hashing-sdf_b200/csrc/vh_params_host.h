// vh_params_host.h — host arithmetic that turns the caller's vh_params and a pose into the kernels' parameter blocks.
// Internal. Kept apart from vh_engine.cu so that the CPU emulation of the kernels (tests/emu, test infrastructure) feeds
// the kernel sources with exactly the constants the engine computes. Must be compiled with -ffp-contract=off (the
// package Makefile passes it to the host compiler): plain float expressions in the reference's order.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>

#include "vh_engine.h"

namespace vh {

// every field of StaticParams that follows from vh_params alone (tuning knobs read from the environment are set by vh_create)
inline void derive_static_params(const vh_params& p, StaticParams& S) {
  S.W = p.width; S.H = p.height; S.fx = p.fx; S.fy = p.fy; S.cx = p.cx; S.cy = p.cy;
  S.min_depth = p.min_depth; S.max_depth = p.max_depth; S.vox_size = p.vox_size; S.trunc = p.trunc_margin;
  S.block_size = (float)VPB * p.vox_size;                                           // tsdf.cu:1326
  S.chunk_size = (float)(p.blocks_per_chunk * VPB) * p.vox_size;                    // tsdf.cu:1271
  S.half_vox = 0.5f * p.vox_size;
  S.stride = p.dda_stride; S.max_steps = p.max_ray_steps; S.bpc = p.blocks_per_chunk;
  const int T = 8;                                                                  // T_PER_BLOCK launch shape, tsdf.cu:2263-2264
  const int gx = (p.width / p.dda_stride + T - 1) / T * T, gy = (p.height / p.dda_stride + T - 1) / T * T;
  S.nrx = std::min(gx, (p.width + p.dda_stride - 1) / p.dda_stride);
  S.nry = std::min(gy, (p.height + p.dda_stride - 1) / p.dda_stride);
  S.use_color = p.use_color ? 1 : 0;
  S.shard_rank = (uint32_t)p.shard_rank; S.shard_count = (uint32_t)p.shard_count;
  S.shard_group = p.shard_group > 0 ? p.shard_group : p.blocks_per_chunk;
  S.byte_bias = 0x4B000000u;
  // approximate-projection error bound (vh_integrate.cu, gate4): 6.9e-7 px per pixel of image extent
  S.round_eps = 7.5e-7f * (float)std::max(p.width, p.height) + 2e-5f;
}

// frame constants: getFrustumCenter / streamInCPU2GPU preamble (tsdf.cu:154-161, :197-206, :300-312)
inline void host_pixel_to_world(const vh_params& P, const float* c2w, int px, int py, float z, float out[3]) {
  const float x = ((float)px - P.cx) * z / P.fx;
  const float y = ((float)py - P.cy) * z / P.fy;
  out[0] = x * c2w[0] + y * c2w[1] + z * c2w[2] + c2w[3];
  out[1] = x * c2w[4] + y * c2w[5] + z * c2w[6] + c2w[7];
  out[2] = x * c2w[8] + y * c2w[9] + z * c2w[10] + c2w[11];
}

// everything in FrameParams except the frame stamp
inline void derive_frame_params(const vh_params& P, const StaticParams& S, const float* c2w, FrameParams& F) {
  memcpy(F.c2w, c2w, sizeof(F.c2w));
  host_pixel_to_world(P, c2w, P.width / 2, P.height / 2, P.max_depth / 2, F.fc);
  const float cs = S.chunk_size;
  const int rng = (int)std::ceil((double)P.chunk_radius / (double)cs);
  const int lo = P.max_chunk_num ? -P.max_chunk_num / 2 : INT32_MIN / 2;
  const int hi = P.max_chunk_num ? P.max_chunk_num / 2 - 1 : INT32_MAX / 2;
  for (int a = 0; a < 3; a++) {
    const int cc = (int)floorf(F.fc[a] / cs);
    F.cstart[a] = std::max(cc - rng, lo);
    F.cend[a] = std::min(cc + rng, hi);
  }
  F.chunk_test_radius = (float)((double)0.5f * (double)P.chunk_radius * (double)sqrtf(3.0f) * 1.1);
}

}  // namespace vh
