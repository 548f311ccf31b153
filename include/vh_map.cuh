// vh_map.cuh — lock-free spatial hash of voxel-block coordinates (device side). Public: include/vhashing.h builds on it.
//
// Replaces the reference's bucketed table with per-bucket spin locks
// (/root/reference/include/vhashing.h:140-239 tryfind/operator[], :395-484 real_insert,
//  include/impl/lockset.h:35-80, include/impl/blockalloc.h:62-103):
//   * one 64-bit word per entry holding the packed block coordinate (3 x 21 bits, biased), claimed with a
//     single atomicCAS(EMPTY -> key) — no locks, no entry chains, no duplicate-insert window;
//   * linear probing over a power-of-two table (the whole table is L2-resident on B200);
//   * value = int32 slot into the voxel-block pools, popped from a free-list stack; the pop is
//     warp-aggregated (one atomicSub per warp per batch, ballot/popc ranks);
//   * the reference's insertion-ordered key_heap / heap_counter (vhashing.h:41-57) is kept: every new key is
//     appended once, also warp-aggregated.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vh {

typedef unsigned long long u64;

constexpr u64 KEY_EMPTY = ~0ull;
constexpr u64 KEY_TOMB = ~0ull - 1;     // erased entry (vh_map_erase only; the TSDF engine never erases)
constexpr int SLOT_UNSET = -1;          // entry claimed, slot not yet published
constexpr int SLOT_POOL_FULL = -2;      // entry claimed while the pool was exhausted: the block has no storage
constexpr int COORD_BIAS = 1 << 20;     // coordinates live in [-2^20, 2^20)

enum MapError { MAP_OK = 0, MAP_TABLE_FULL = 1, MAP_POOL_FULL = 2, MAP_KEY_RANGE = 4 };

struct MapView {
  u64* keys;            // [capacity]   packed key or KEY_EMPTY / KEY_TOMB
  int* slots;           // [capacity]   pool slot of the entry's block (-1 until assigned)
  uint32_t mask;        // capacity - 1
  int* free_list;       // [num_blocks] stack of free pool slots
  int* free_top;        // number of free slots left on the stack
  u64* key_heap;        // [num_blocks] keys in insertion order
  int* heap_counter;    // number of keys appended
  int* error_flag;      // MapError bits, sticky
  int num_blocks;
};

__host__ __device__ __forceinline__ u64 pack_key(int x, int y, int z) {
  return ((u64)(uint32_t)(x + COORD_BIAS) << 42) | ((u64)(uint32_t)(y + COORD_BIAS) << 21) | (u64)(uint32_t)(z + COORD_BIAS);
}
__host__ __device__ __forceinline__ bool key_in_range(int x, int y, int z) {
  return ((uint32_t)(x + COORD_BIAS) | (uint32_t)(y + COORD_BIAS) | (uint32_t)(z + COORD_BIAS)) < (1u << 21);
}
__host__ __device__ __forceinline__ void unpack_key(u64 k, int& x, int& y, int& z) {
  x = (int)((k >> 42) & 0x1FFFFF) - COORD_BIAS;
  y = (int)((k >> 21) & 0x1FFFFF) - COORD_BIAS;
  z = (int)(k & 0x1FFFFF) - COORD_BIAS;
}
// murmur3 finaliser: neighbouring blocks land in unrelated probe sequences
__host__ __device__ __forceinline__ uint32_t hash_key(u64 k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k;
}
// a different mix decides which GPU owns a block, so shards do not correlate with probe positions
__host__ __device__ __forceinline__ uint32_t owner_of_key(u64 k, uint32_t shard_count) {
  k *= 0x9E3779B97F4A7C15ull; k ^= k >> 29; k *= 0xBF58476D1CE4E5B9ull; k ^= k >> 32;
  return (uint32_t)(k % shard_count);
}
// Ownership is decided per cube of `group` blocks per axis (group = 1: per block; group = 8: per reference chunk). A
// block and its +x/+y/+z neighbours then share an owner unless they straddle a cube face: with group = 8 four out of
// five marching-cubes neighbour look-ups stay on the local GPU.
__host__ __device__ __forceinline__ int floor_div_int(int a, int b) { const int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__host__ __device__ __forceinline__ uint32_t owner_of_block(int x, int y, int z, uint32_t shard_count, int group) {
  if (group > 1) { x = floor_div_int(x, group); y = floor_div_int(y, group); z = floor_div_int(z, group); }
  return owner_of_key(pack_key(x, y, z), shard_count);
}

#ifdef __CUDACC__
// read-only lookup in a table given by its key array (own or a peer GPU's, mapped): entry index or -1
__device__ __forceinline__ int map_find_in(const u64* __restrict__ keys, uint32_t mask, u64 key) {
  uint32_t h = hash_key(key) & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    const u64 cur = __ldcg(&keys[h]);
    if (cur == key) return (int)h;
    if (cur == KEY_EMPTY) return -1;
    h = (h + 1) & mask;
  }
  return -1;
}
__device__ __forceinline__ int map_find(const MapView& m, u64 key) { return map_find_in(m.keys, m.mask, key); }

// insert-if-absent: entry index (>= 0) and whether THIS call claimed the entry; -1 when the table is full
__device__ __forceinline__ int map_claim(const MapView& m, u64 key, bool& claimed) {
  claimed = false;
  uint32_t h = hash_key(key) & m.mask;
  for (uint32_t probe = 0; probe <= m.mask; ++probe) {
    u64 cur = __ldcg(&m.keys[h]);
    if (cur == key) return (int)h;
    if (cur == KEY_EMPTY) {
      cur = atomicCAS(&m.keys[h], KEY_EMPTY, key);
      if (cur == KEY_EMPTY) { claimed = true; return (int)h; }
      if (cur == key) return (int)h;
    }
    h = (h + 1) & m.mask;
  }
  atomicOr(m.error_flag, MAP_TABLE_FULL);
  return -1;
}

// Warp-aggregated pool allocation + key_heap append for the lanes that just claimed an entry.
// Must be called by all lanes named in `active` (converged). One atomicSub and one atomicAdd per warp.
// Pool exhausted: the lanes that got no slot publish SLOT_POOL_FULL (what vhashing.h's wait_slot expects; -1 means "claimed,
// not yet published" and would make a reader spin), their share of the pop is handed back so that free_top never stays
// negative (a later push — erase, evict — must land inside the stack; the reference clamps with atomicMax(link_head, -1),
// blockalloc.h:62-86), and heap_counter advances only by the slots actually granted, so key_heap has no holes.
__device__ __forceinline__ void map_assign_slots(const MapView& m, unsigned active, bool claimed, int entry, u64 key) {
  const unsigned newmask = __ballot_sync(active, claimed);
  if (newmask == 0) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(newmask) - 1;
  const int n = __popc(newmask);
  int top = 0, heap_base = 0;
  if (lane == leader) {
    top = atomicSub(m.free_top, n);
    const int granted = top >= n ? n : (top > 0 ? top : 0);
    if (granted < n) atomicAdd(m.free_top, n - granted);
    if (granted > 0) heap_base = atomicAdd(m.heap_counter, granted);
  }
  top = __shfl_sync(active, top, leader);
  heap_base = __shfl_sync(active, heap_base, leader);
  if (claimed) {
    const int rank = __popc(newmask & ((1u << lane) - 1));
    const int idx = top - 1 - rank;
    if (idx >= 0) {
      m.slots[entry] = m.free_list[idx];
      if (heap_base + rank < m.num_blocks) m.key_heap[heap_base + rank] = key;
    } else {
      m.slots[entry] = SLOT_POOL_FULL;
      atomicOr(m.error_flag, MAP_POOL_FULL);
    }
  }
}

// push a slot back on the free-list stack (erase / evict). A position outside the stack can only come from a pop that is
// handing back its share at this very moment (see map_assign_slots): the slot is dropped rather than written out of bounds.
__device__ __forceinline__ void map_release_slot(const MapView& m, int slot) {
  const int pos = atomicAdd(m.free_top, 1);
  if (pos >= 0 && pos < m.num_blocks) m.free_list[pos] = slot; else atomicOr(m.error_flag, MAP_POOL_FULL);
}
#endif  // __CUDACC__

}  // namespace vh
