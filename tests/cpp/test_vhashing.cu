// test_vhashing.cu — vhashing::HashTable drop-in (include/vhashing.h) used the way the reference's kernels use it
// (/root/reference/src/tsdf.cu:208-216 insert through operator[], :459-467 key_heap walk, :2164 find != end()).
#include <tsdf.cuh>
#include <vhashing.h>

#include <cstdio>
#include <set>
#include <string>
#include <tuple>
#include <vector>

struct Payload {
  int hits;
  float tag;
  __host__ __device__ Payload() : hits(0), tag(-1.0f) {}
};
typedef vhashing::HashTable<int3, Payload, ark::BlockHasher, ark::BlockEqual, vhashing::device_memspace> Table;
typedef vhashing::HashTableBase<int3, Payload, ark::BlockHasher, ark::BlockEqual> TableBase;

__host__ __device__ inline int3 key_of(int i, int uniq) {
  const int j = i % uniq;
  return make_int3(j % 37 - 18, (j / 37) % 41 - 20, j / (37 * 41) - 3);
}

// many threads, heavy key duplication: operator[] must hand every thread of a key the same value
__global__ void insert_kernel(TableBase t, int n, int uniq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Payload& p = t[key_of(i, uniq)];
  atomicAdd(&p.hits, 1);
}
__global__ void find_kernel(TableBase t, int uniq, int* present, int* absent_found, long long* hit_sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= uniq) return;
  const int3 k = key_of(i, uniq);
  auto it = t.find(k);
  if (it != t.end()) {
    atomicAdd(present, 1);
    if (it->key.x != k.x || it->key.y != k.y || it->key.z != k.z || it->offset != 0) atomicAdd(absent_found, 1000);
    atomicAdd((unsigned long long*)hit_sum, (unsigned long long)t[*it].hits);
    const TableBase& ct = t;
    if (ct[k].hits != t[*it].hits) atomicAdd(absent_found, 1000);
  }
  if (t.find(make_int3(k.x + 1000, k.y, k.z)) != t.end()) atomicAdd(absent_found, 1);
}
__global__ void heap_kernel(TableBase t, int* bad) {      // getMapValueKernel pattern: walk key_heap[0 .. heap_counter)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *t.heap_counter) return;
  const int3 k = t.key_heap[i];
  if (t.find(k) == t.end()) atomicAdd(bad, 1);
}
__global__ void tryinsert_erase_kernel(TableBase t, int* out) {
  if (threadIdx.x || blockIdx.x) return;
  Payload p; p.hits = 7; p.tag = 3.5f;
  const int3 k = make_int3(500, -500, 123);
  auto a = t.tryinsert(k, p);
  Payload q; q.hits = 9;
  auto b = t.tryinsert(k, q);                       // already there: same entry, value untouched
  out[0] = (a != t.end()) && (a == b) && t[k].hits == 7 && t[k].tag == 3.5f;
  out[1] = t.erase(k);
  out[2] = t.find(k) == t.end();
  out[3] = t.erase(k) == 0;
  t[k].hits = 1;                                    // re-insert after erase: default-constructed value
  out[4] = t[k].tag == -1.0f && t[k].hits == 1;
}
struct TagOp { __device__ void operator()(const int3& k, Payload& p) const { p.tag = (float)(k.x + k.y + k.z); } };
struct EvenX { __device__ bool operator()(const int3& k, const Payload&) const { return (k.x & 1) == 0; } };
__global__ void check_tags(const TableBase::HashEntry* e, int n, TableBase t, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if ((e[i].key.x & 1) != 0 || t[e[i]].tag != (float)(e[i].key.x + e[i].key.y + e[i].key.z)) atomicAdd(bad, 1);
}

// after a bulk allocation that overflowed the pool: device-side access to every requested key terminates (keys that got no
// storage read as absent-with-error, not as "slot not yet published"), erase pushes stay inside the free-list, freed slots are reusable
__global__ void touch_after_overflow(TableBase t, int n, int* reached, int* stored) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int3 k = make_int3(7000 + i, -3, i & 7);
  const TableBase& ct = t;
  const Payload& p = ct[k];                                        // const access: wait_slot must not spin on a pool-full entry
  if (p.hits >= 0) atomicAdd(reached, 1);
  if (t.find(k) != t.end() && t.find(k)->block_index >= 0) atomicAdd(stored, 1);
}
__global__ void erase_some(TableBase t, int n, int* erased) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && t.erase(make_int3(7000 + i, -3, i & 7))) atomicAdd(erased, 1);
}

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
  const int N = 1 << 18, UNIQ = 5000;
  Table table(4096, 4, 8192, make_int3(999999, 999999, 999999));      // the reference's ctor spelling (tsdf.cu:1488)
  int *d = nullptr; long long* d_sum = nullptr;
  cudaMalloc(&d, 8 * sizeof(int)); cudaMalloc(&d_sum, sizeof(long long));
  cudaMemset(d, 0, 8 * sizeof(int)); cudaMemset(d_sum, 0, sizeof(long long));

  insert_kernel<<<N / 256, 256>>>(table, N, UNIQ);
  find_kernel<<<(UNIQ + 255) / 256, 256>>>(table, UNIQ, d, d + 1, d_sum);
  heap_kernel<<<(8192 + 255) / 256, 256>>>(table, d + 2);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  int h[8]; long long sum = 0;
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&sum, d_sum, sizeof(sum), cudaMemcpyDeviceToHost);
  table.check();
  REQUIRE(h[0] == UNIQ);            // every key present
  REQUIRE(h[1] == 0);               // no absent key found, entries consistent
  REQUIRE(sum == N);                // no duplicate inserts: all increments landed in one value per key
  REQUIRE(h[2] == 0);
  REQUIRE(table.size() == UNIQ);
  int heap = 0; cudaMemcpy(&heap, table.heap_counter, sizeof(int), cudaMemcpyDeviceToHost);
  REQUIRE(heap == UNIQ);
  std::set<std::tuple<int, int, int>> want, got;
  for (int i = 0; i < UNIQ; i++) { const int3 k = key_of(i, UNIQ); want.insert({k.x, k.y, k.z}); }
  for (const int3& k : table.Keys()) got.insert({k.x, k.y, k.z});
  REQUIRE(want == got);

  // host bulk allocation with duplicates and already-present keys
  std::vector<int3> more;
  for (int i = 0; i < 300; i++) more.push_back(make_int3(2000 + i % 100, 7, -9));
  more.push_back(key_of(3, UNIQ));
  table.AllocKeys(more);
  REQUIRE(table.size() == UNIQ + 100);

  tryinsert_erase_kernel<<<1, 32>>>(table, d + 3);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 3; i < 8; i++) REQUIRE(h[i] == 1);

  table.Apply(TagOp());
  auto filtered = table.Filter(EvenX());
  REQUIRE(filtered.second > 0 && filtered.second < table.size());
  cudaMemset(d, 0, sizeof(int));
  check_tags<<<(filtered.second + 255) / 256, 256>>>(filtered.first.get(), filtered.second, table, d);
  cudaMemcpy(h, d, sizeof(int), cudaMemcpyDeviceToHost);
  REQUIRE(h[0] == 0);
  auto all = table.Filter();
  REQUIRE(all.second == table.size());

  // exhaustion is an error, not a hang (reference: operator[] spins forever, vhashing.h:216-231)
  Table small(64, 4, 100, make_int3(999999, 999999, 999999));
  insert_kernel<<<4, 256>>>(small, 1024, 1000);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  bool threw = false;
  try { small.check(); } catch (const char* m) { threw = std::string(m) == "out of block memory"; }
  REQUIRE(threw);

  // bulk allocation past the pool: the host call raises, and the table stays usable from device code
  Table tiny(256, 4, 50, make_int3(999999, 999999, 999999));
  std::vector<int3> over;
  for (int i = 0; i < 120; i++) over.push_back(make_int3(7000 + i, -3, i & 7));
  threw = false;
  try { tiny.AllocKeys(over); } catch (const char* m) { threw = std::string(m) == "out of block memory"; }
  REQUIRE(threw);
  REQUIRE(tiny.size() == 50);                                        // key_heap lists exactly the blocks that got storage
  cudaMemset(d, 0, 8 * sizeof(int));
  touch_after_overflow<<<1, 128>>>(tiny, 120, d, d + 1);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);                   // no spin on SLOT_POOL_FULL entries
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  REQUIRE(h[0] == 120 && h[1] == 50);
  erase_some<<<1, 128>>>(tiny, 120, d + 2);
  REQUIRE(cudaDeviceSynchronize() == cudaSuccess);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  REQUIRE(h[2] == 120);                                              // every claimed entry goes, 50 slots return to the stack (inside its bounds)
  int top = 0; cudaMemcpy(&top, tiny.view.free_top, sizeof(int), cudaMemcpyDeviceToHost);
  REQUIRE(top == 50);
  std::vector<int3> again;
  for (int i = 0; i < 50; i++) again.push_back(make_int3(-40 - i, 2, 2));
  REQUIRE(tiny.size() == 0);
  try { tiny.AllocKeys(again); } catch (const char*) {}              // (the error flag is sticky: the call still reports the earlier overflow)
  REQUIRE(tiny.size() == 50);                                        // the freed slots were handed out again
  std::printf("VHASHING_OK\n");
  return 0;
}
