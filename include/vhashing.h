// vhashing.h — drop-in for the reference's vhashing::HashTable / HashTableBase
// (/root/reference/include/vhashing.h:35-826, impl/memspace.h:8-10, impl/query.h) on top of the lock-free
// B200 block hash of this repo (include/vh_map.cuh, C ABI vh_map_* in include/vh_c.h).
//
// Same spelling for callers:
//     vhashing::HashTable<int3, Value, Hash, Equal, vhashing::device_memspace> table(buckets, entries, blocks, emptyKey);
//     kernel<<<g, b>>>(table /* sliced to HashTableBase, passed by value */);
//     __device__: table.find(k) != table.end(), table[k], it->key / it->block_index, table.key_heap[i], *table.heap_counter
//     host:       table.AllocKeys(std::vector<Key>), table.Filter(pred), table.Apply(op)
//
// What is different underneath (and why the results of a race-free program are the same):
//   * no bucket locks (reference: impl/lockset.h): an entry is claimed by one 64-bit atomicCAS on the packed key;
//   * no entry chains: linear probing in a power-of-two table; `offset` of an entry is always 0;
//   * value slots come from a free-list stack; `block_index` is the slot, `alloc[slot]` the value;
//   * insert-if-absent is exact: two threads inserting the same key get the same slot (the reference's operator[]
//     can insert a key twice);
//   * a full table or an exhausted pool raises a sticky error flag and returns end() / a scratch value instead of
//     spinning forever (reference: vhashing.h:216-231) — HashTable::check() turns the flag into the reference's
//     host exceptions ("Error here!", "out of block memory").
// Keys are block coordinates: any type with int members x, y, z in [-2^20, 2^20). Only device_memspace tables exist
// (this engine has no CPU path); the host_memspace / std_memspace tags are declared so that code naming them
// still parses, and instantiating a table with them is a compile-time error.
//
// Device-side members need nvcc (they are guarded by __CUDACC__); the host-side class works from plain C++.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <new>
#include <type_traits>
#include <utility>
#include <vector>

#include "vh_c.h"
#include "vh_map.cuh"

namespace vhashing {

struct device_memspace {};
struct host_memspace {};
typedef host_memspace std_memspace;

template <typename Key>
struct HashEntryBase {
  Key key;
  int32_t offset;        // always 0: there are no overflow chains
  int32_t block_index;   // value slot
};

namespace detail {

struct AlwaysTrue {
  template <typename Key, typename T>
  __host__ __device__ bool operator()(const Key&, const T&) const { return true; }
};

template <class Key>
__host__ __device__ inline vh::u64 encode(const Key& k) { return vh::pack_key(k.x, k.y, k.z); }
template <class Key>
__host__ __device__ inline Key decode(vh::u64 p) {
  Key k{};
  int x, y, z;
  vh::unpack_key(p, x, y, z);
  k.x = x; k.y = y; k.z = z;
  return k;
}

template <class Value>
struct SlotPool {            // the reference's BlockAllocBase seen from device code: alloc[slot], alloc.data
  Value* data;
  int num_elems;
  __host__ __device__ Value& operator[](int slot) const { return data[slot]; }
};

template <class Key>
struct KeyHeapView {          // key_heap[i]: the i-th inserted key (stored packed; decoded on access)
  const vh::u64* packed;
  __host__ __device__ Key operator[](int i) const { return decode<Key>(packed[i]); }
};

struct memspace_free {
  void operator()(void* p) const { if (p) cudaFree(p); }
};

}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------
// The by-value device view (reference: HashTableBase, vhashing.h:31-603).
template <class Key, class Value, class Hash, class Equal>
struct HashTableBase {
  typedef Key KeyType;
  typedef Value ValueType;
  typedef HashEntryBase<Key> HashEntry;
  typedef detail::SlotPool<Value> BlockAlloc;

  int num_buckets;
  int entries_per_bucket;
  uint32_t num_entries;            // table capacity (power of two >= num_buckets * entries_per_bucket)
  int* heap_counter;               // number of keys inserted so far
  Key emptyKey;
  Hash hasher;                     // kept for source compatibility; probing uses the table's own 64-bit mix
  Equal isequal;
  detail::KeyHeapView<Key> key_heap;
  BlockAlloc alloc;
  vh::MapView view;
  Value* scratch;                  // returned by operator[] after an error so that callers never write through null

  __host__ __device__ Key EmptyKey() const { return emptyKey; }

  struct iterator {
    const HashTableBase* bm;
    int32_t offset;                // entry index; -1 = end(), -2 = fail()
    struct arrow { HashEntry he; __host__ __device__ const HashEntry* operator->() const { return &he; } };
#ifdef __CUDACC__
    __device__ HashEntry operator*() const {
      HashEntry he;
      he.key = detail::decode<Key>(bm->view.keys[offset]); he.offset = 0; he.block_index = bm->view.slots[offset];
      return he;
    }
    __device__ arrow operator->() const { return arrow{**this}; }
#endif
    __host__ __device__ bool operator==(const iterator& b) const { return offset == b.offset; }
    __host__ __device__ bool operator!=(const iterator& b) const { return offset != b.offset; }
  };
  __host__ __device__ iterator end() const { return iterator{this, -1}; }
  __host__ __device__ iterator fail() const { return iterator{this, -2}; }

#ifdef __CUDACC__
  __device__ bool IsEmpty(int32_t off) const { return view.keys[off] == vh::KEY_EMPTY || view.keys[off] == vh::KEY_TOMB; }
  __device__ void clearheap() { atomicExch(heap_counter, 0); }

  // read-only lookup, no atomics (reference: find = tryfind(k, readonly), vhashing.h:140-142)
  __device__ iterator find(const Key& k) const { return iterator{this, vh::map_find(view, detail::encode(k))}; }

  __device__ Value& operator[](const HashEntry& he) const { return alloc.data[he.block_index]; }

  // read-only access; the key must be present (reference asserts, vhashing.h:129-137)
  __device__ Value& operator[](const Key& k) const {
    const int e = vh::map_find(view, detail::encode(k));
    const int s = e >= 0 ? wait_slot(e) : -1;
    if (s < 0) { atomicOr(view.error_flag, vh::MAP_KEY_RANGE); return *scratch; }
    return alloc.data[s];
  }

  // access, inserting a default-constructed value when the key is absent (reference: vhashing.h:206-239)
  __device__ Value& operator[](const Key& k) {
    const int s = insert_slot(k, nullptr);
    return s >= 0 ? alloc.data[s] : *scratch;
  }

  // insert-if-absent with a value; returns the entry (existing or new) or end() when the table/pool is full
  __device__ iterator tryinsert(const Key& k, const Value& v) {
    int entry = -1;
    const int s = insert_slot(k, &v, &entry);
    return s >= 0 ? iterator{this, entry} : end();
  }
  __device__ iterator insert(const Key& k, const Value& v) { return tryinsert(k, v); }

  // remove the key; must not race with inserts of the same key (the reference serialises them on the bucket lock)
  __device__ int erase(const Key& k) {
    const vh::u64 p = detail::encode(k);
    const int e = vh::map_find(view, p);
    if (e < 0) return 0;
    const int s = wait_slot(e);
    if (atomicCAS(&view.keys[e], p, vh::KEY_TOMB) != p) return 0;
    view.slots[e] = -1;
    if (s >= 0) vh::map_release_slot(view, s);
    return 1;
  }

 private:
  __device__ int wait_slot(int e) const {
    int s;
    while ((s = *reinterpret_cast<volatile int*>(&view.slots[e])) == vh::SLOT_UNSET) {}   // claimed, slot not yet published
    return s;                                                                 // >= 0, or -2 = pool was exhausted
  }
  __device__ int insert_slot(const Key& k, const Value* v, int* entry_out = nullptr) {
    if (!vh::key_in_range(k.x, k.y, k.z)) { atomicOr(view.error_flag, vh::MAP_KEY_RANGE); return -1; }
    const vh::u64 p = detail::encode(k);
    bool claimed = false;
    const int e = vh::map_claim(view, p, claimed);
    if (e < 0) return -1;
    if (entry_out) *entry_out = e;
    if (!claimed) return wait_slot(e);
    int slot = vh::SLOT_POOL_FULL;
    const int top = atomicSub(view.free_top, 1);
    if (top > 0) {
      slot = view.free_list[top - 1];
      const int pos = atomicAdd(view.heap_counter, 1);
      if (pos < view.num_blocks) view.key_heap[pos] = p;
      if (v) new (&alloc.data[slot]) Value(*v); else new (&alloc.data[slot]) Value();
      __threadfence();
    } else {
      atomicAdd(view.free_top, 1);
      atomicOr(view.error_flag, vh::MAP_POOL_FULL);
    }
    atomicExch(&view.slots[e], slot);
    return slot;
  }
#endif  // __CUDACC__
};

// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace detail {
template <class Value>
__global__ void construct_values(Value* data, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) new (&data[i]) Value();
}
template <class Base, class Fil>
__global__ void filter_entries(Base t, Fil f, typename Base::HashEntry* out, int* n_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.num_entries || t.IsEmpty((int32_t)i)) return;
  typename Base::HashEntry he;
  he.key = decode<typename Base::KeyType>(t.view.keys[i]); he.offset = 0; he.block_index = t.view.slots[i];
  if (he.block_index >= 0 && f(he.key, t.alloc.data[he.block_index])) out[atomicAdd(n_out, 1)] = he;
}
template <class Base, class Op>
__global__ void apply_entries(Base t, Op op) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.num_entries || t.IsEmpty((int32_t)i)) return;
  const int s = t.view.slots[i];
  if (s >= 0) op(decode<typename Base::KeyType>(t.view.keys[i]), t.alloc.data[s]);
}
}  // namespace detail
#endif

// The owning host class (reference: HashTable, vhashing.h:617-826). Copying it to a kernel argument slices it to
// HashTableBase exactly as with the reference.
template <class Key, class Value, class Hash, class Equal, class memspace = device_memspace>
class HashTable : public HashTableBase<Key, Value, Hash, Equal> {
  static_assert(std::is_same<memspace, device_memspace>::value,
                "this engine keeps its tables in HBM only: use vhashing::device_memspace (there is no CPU fallback)");

 public:
  typedef HashTableBase<Key, Value, Hash, Equal> parent_type;
  typedef std::unique_ptr<typename parent_type::HashEntry, detail::memspace_free> HashEntriesPtr;

  HashTable(int num_buckets, int entries_per_bucket, int num_blocks, Key emptyKey, Hash hasher = Hash(), Equal equals = Equal(), int device = 0) {
    this->num_buckets = num_buckets; this->entries_per_bucket = entries_per_bucket;
    this->emptyKey = emptyKey; this->hasher = hasher; this->isequal = equals;
    if (vh_map_create(num_buckets, entries_per_bucket, num_blocks, device, &map_) != VH_OK) throw "Error here!";   // vhashing.h:106
    vh_map_view v;
    vh_map_get_view(map_, &v);
    this->view.keys = v.keys; this->view.slots = v.slots; this->view.mask = v.capacity_mask; this->view.free_list = v.free_list;
    this->view.free_top = v.free_top; this->view.key_heap = v.key_heap; this->view.heap_counter = v.heap_counter;
    this->view.error_flag = v.error_flag; this->view.num_blocks = v.num_blocks;
    this->num_entries = v.capacity_mask + 1u;
    this->heap_counter = v.heap_counter;
    this->key_heap.packed = v.key_heap;
    Value* data = nullptr;
    if (cudaMalloc((void**)&data, ((size_t)num_blocks + 1) * sizeof(Value)) != cudaSuccess) { vh_map_destroy(map_); map_ = nullptr; throw "CUDA Error"; }
    this->alloc.data = data; this->alloc.num_elems = num_blocks;
    this->scratch = data + num_blocks;
#ifdef __CUDACC__
    detail::construct_values<Value><<<(num_blocks + 256) / 256, 256>>>(data, num_blocks + 1);   // data_shared(num_blocks + 1), vhashing.h:653
    if (cudaDeviceSynchronize() != cudaSuccess) throw "CUDA Error";
#else
    cudaMemset(data, 0, ((size_t)num_blocks + 1) * sizeof(Value));
#endif
  }
  HashTable(const HashTable&) = delete;            // one owner per table; pass parent_type to kernels
  HashTable& operator=(const HashTable&) = delete;
  HashTable(HashTable&& o) : parent_type(o), map_(o.map_) { o.map_ = nullptr; o.alloc.data = nullptr; }
  ~HashTable() {
    if (this->alloc.data) cudaFree(this->alloc.data);
    if (map_) vh_map_destroy(map_);
  }

  // sticky device-side errors as the reference's host exceptions
  void check() const {
    int f = 0;
    if (cudaMemcpy(&f, this->view.error_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) throw "CUDA Error";
    if (f & vh::MAP_POOL_FULL) throw "out of block memory";    // impl/blockalloc.h:51
    if (f) throw "Error here!";                                // vhashing.h:106
  }
  int size() const { int n = 0; vh_map_size(map_, &n); return n; }
  vh_map* handle() const { return map_; }

  // bulk allocation (reference: AllocKeys / AllocKeysNoDups, vhashing.h:531-603); duplicates are fine
  void AllocKeys(const std::vector<Key>& keys) {
    std::vector<int32_t> xyz(keys.size() * 3);
    for (size_t i = 0; i < keys.size(); i++) { xyz[3 * i] = keys[i].x; xyz[3 * i + 1] = keys[i].y; xyz[3 * i + 2] = keys[i].z; }
    const int rc = vh_map_insert(map_, xyz.data(), (int)keys.size(), nullptr);
    if (rc == VH_ERR_POOL_FULL) throw "out of block memory";
    if (rc != VH_OK) throw "Error here!";
  }
  void AllocKeysNoDups(const std::vector<Key>& keys, bool /*retry*/ = false) { AllocKeys(keys); }

  // keys in insertion order (reference callers read key_heap[0 .. *heap_counter))
  std::vector<Key> Keys() const {
    int n = 0;
    vh_map_keys(map_, nullptr, 0, &n);
    std::vector<int32_t> xyz((size_t)(n > 0 ? n : 1) * 3);
    vh_map_keys(map_, xyz.data(), n, &n);
    std::vector<Key> out((size_t)n);
    for (int i = 0; i < n; i++) { out[i] = Key{}; out[i].x = xyz[3 * i]; out[i].y = xyz[3 * i + 1]; out[i].z = xyz[3 * i + 2]; }
    return out;
  }

#ifdef __CUDACC__
  // [entries, n]: device array of the entries with filter(key, value) == true (reference: Filter, vhashing.h:790-813)
  template <class Fil = detail::AlwaysTrue>
  std::pair<HashEntriesPtr, int> Filter(Fil filter = Fil()) const {
    typename parent_type::HashEntry* d = nullptr;
    int* d_n = nullptr;
    int n = 0;
    if (cudaMalloc((void**)&d, sizeof(typename parent_type::HashEntry) * (size_t)this->alloc.num_elems) != cudaSuccess ||
        cudaMalloc((void**)&d_n, sizeof(int)) != cudaSuccess) throw "CUDA Error";
    cudaMemset(d_n, 0, sizeof(int));
    detail::filter_entries<parent_type, Fil><<<(this->num_entries + 255) / 256, 256>>>(*this, filter, d, d_n);
    cudaMemcpy(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_n);
    return std::make_pair(HashEntriesPtr(d), n);
  }
  // op(key, value&) on every live entry (reference: Apply, vhashing.h:818-826)
  template <class Op>
  void Apply(Op op = Op()) {
    detail::apply_entries<parent_type, Op><<<(this->num_entries + 255) / 256, 256>>>(*this, op);
    if (cudaDeviceSynchronize() != cudaSuccess) throw "CUDA Error";
  }
#endif

 private:
  vh_map* map_ = nullptr;
};

}  // namespace vhashing
