/* Headless C-ABI driver around the reference's own GpuTsdfGenerator, compiled for the CPU
 * through emu_shim.h (test scaffolding, NOT product code; built only into oracle/_ref/).
 * It calls the reference's processFrame / SavePLY unchanged and reads results back out of
 * the reference's own host chunk store (h_chunks, tsdf.cuh:167-218).
 */
#include "tsdf.cuh"
#include <vector>
#include <cstdint>
#include <cstring>

using namespace ark;

/* filled by the one-line hook that build_ref.sh inserts before streamOutGPU2CPU() (tsdf.cu:1585) */
static std::vector<int3> g_last_visible;
namespace ark { void emu_hook_visible(const int3* key_heap, int n) { g_last_visible.assign(key_heap, key_heap + n); } }

struct RefHandle { GpuTsdfGenerator* gen; };

static inline int chunk_linear(int x, int y, int z) {
  const int h = MAX_CHUNK_NUM / 2;
  return ((x + h) * MAX_CHUNK_NUM + (y + h)) * MAX_CHUNK_NUM + (z + h);
}
static inline int fdiv(int a, int b) { int q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }

extern "C" {

int ref_voxels_per_block() { return VOXEL_PER_BLOCK; }

void* ref_create(int W, int H, float fx, float fy, float cx, float cy, float max_depth, float vox_size, float trunc) {
  RefHandle* h = new RefHandle;
  /* tiny Voxel.Dim keeps the unused legacy dense grids small (tsdf.cu:1343-1351) */
  h->gen = new GpuTsdfGenerator(W, H, fx, fy, cx, cy, max_depth, 0.f, 0.f, 0.f, vox_size, trunc, 8, 8, 8);
  return h;
}

void ref_process_frame(void* hv, float* depth, unsigned char* rgb, float* c2w) {
  ((RefHandle*)hv)->gen->processFrame(depth, rgb, c2w);
}

int ref_last_visible_count(void* hv) { return (int)((RefHandle*)hv)->gen->h_heapBlockCounter; }
int ref_last_streamed_blocks(void* hv) { return (int)((RefHandle*)hv)->gen->h_inChunkCounter; }

/* visible key list of the last frame, in the (sequential) insertion order; returns count */
int ref_last_visible_keys(void*, int* out_xyz, int cap) {
  int n = (int)g_last_visible.size();
  for (int i = 0; i < n && i < cap; i++) { out_xyz[3*i] = g_last_visible[i].x; out_xyz[3*i+1] = g_last_visible[i].y; out_xyz[3*i+2] = g_last_visible[i].z; }
  return n;
}

/* copy one stored block out of the host chunk store; returns 0 if its chunk does not exist */
int ref_get_block(void* hv, int bx, int by, int bz, float* sdf, float* weight, unsigned char* rgb) {
  GpuTsdfGenerator* g = ((RefHandle*)hv)->gen;
  int cx = fdiv(bx, BLOCK_PER_CHUNK), cy = fdiv(by, BLOCK_PER_CHUNK), cz = fdiv(bz, BLOCK_PER_CHUNK);
  const int hh = MAX_CHUNK_NUM / 2;
  if (cx < -hh || cx >= hh || cy < -hh || cy >= hh || cz < -hh || cz >= hh) return 0;
  Chunk& c = g->h_chunks[chunk_linear(cx, cy, cz)];
  if (c.blocks == nullptr) return 0;
  int lx = bx - cx * BLOCK_PER_CHUNK, ly = by - cy * BLOCK_PER_CHUNK, lz = bz - cz * BLOCK_PER_CHUNK;
  VoxelBlock& vb = c.blocks[(lx * BLOCK_PER_CHUNK + ly) * BLOCK_PER_CHUNK + lz];
  const int n = VOXEL_PER_BLOCK * VOXEL_PER_BLOCK * VOXEL_PER_BLOCK;
  for (int i = 0; i < n; i++) {
    sdf[i] = vb.voxels[i].sdf; weight[i] = vb.voxels[i].weight;
    if (rgb) { rgb[3*i] = vb.voxels[i].sdf_color[0]; rgb[3*i+1] = vb.voxels[i].sdf_color[1]; rgb[3*i+2] = vb.voxels[i].sdf_color[2]; }
  }
  return 1;
}

/* checksum over every stored voxel of every existing chunk */
void ref_voxel_checksum(void* hv, double* sum_sdf, double* sum_w, long long* n_observed, long long* n_negative) {
  GpuTsdfGenerator* g = ((RefHandle*)hv)->gen;
  const int NC = MAX_CHUNK_NUM * MAX_CHUNK_NUM * MAX_CHUNK_NUM;
  const int nb = BLOCK_PER_CHUNK * BLOCK_PER_CHUNK * BLOCK_PER_CHUNK;
  const int nv = VOXEL_PER_BLOCK * VOXEL_PER_BLOCK * VOXEL_PER_BLOCK;
  double ss = 0, sw = 0; long long no = 0, nn = 0;
  for (int c = 0; c < NC; c++) {
    Chunk& ck = g->h_chunks[c];
    if (!ck.blocks) continue;
    for (int b = 0; b < nb; b++) for (int v = 0; v < nv; v++) {
      const Voxel& vx = ck.blocks[b].voxels[v];
      ss += vx.sdf; sw += vx.weight; no += (vx.weight > 0); nn += (vx.sdf < 0);
    }
  }
  *sum_sdf = ss; *sum_w = sw; *n_observed = no; *n_negative = nn;
}

/* Valid triangles in the order tsdf2mesh walks them (tsdf.cu:1786-1806): chunks x,y,z ascending,
 * slot ascending. Coordinates are raw voxel-index units (before the vox_size scale at :1815).
 * out_xyz: 9 floats per triangle, out_rgb: 9 bytes per triangle (either may be NULL). Returns the count. */
long long ref_triangles(void* hv, float* out_xyz, unsigned char* out_rgb, long long cap) {
  GpuTsdfGenerator* g = ((RefHandle*)hv)->gen;
  const int hh = MAX_CHUNK_NUM / 2;
  const int nb = BLOCK_PER_CHUNK * BLOCK_PER_CHUNK * BLOCK_PER_CHUNK;
  const long long slots = (long long)nb * VOXEL_PER_BLOCK * VOXEL_PER_BLOCK * VOXEL_PER_BLOCK * 5;
  long long n = 0;
  for (int x = -hh; x < hh; x++) for (int y = -hh; y < hh; y++) for (int z = -hh; z < hh; z++) {
    Chunk& ck = g->h_chunks[chunk_linear(x, y, z)];
    if (!ck.isOccupied || !ck.tri_) continue;   /* tsdf2mesh releases unoccupied chunks first (:1790) */
    for (long long s = 0; s < slots; s++) {
      const Triangle& t = ck.tri_[s];
      if (!t.valid) continue;
      if (n < cap) {
        for (int j = 0; j < 3; j++) {
          if (out_xyz) { out_xyz[9*n+3*j] = t.p[j].x; out_xyz[9*n+3*j+1] = t.p[j].y; out_xyz[9*n+3*j+2] = t.p[j].z; }
          if (out_rgb) { out_rgb[9*n+3*j] = t.p[j].r; out_rgb[9*n+3*j+1] = t.p[j].g; out_rgb[9*n+3*j+2] = t.p[j].b; }
        }
      }
      n++;
    }
  }
  return n;
}

void ref_save_ply(void* hv, const char* path) { ((RefHandle*)hv)->gen->SavePLY(std::string(path)); }

/* direct access to the reference's __host__ __device__ math for known-answer tests */
}
namespace ark {
  void frame2cam(int* pt_pix, float pt_cam_z, float* pt_cam, float* K_);
  void cam2frame(float* pt_cam, int* pt_pix, float* K);
  void base2cam(float* pt_base, float* pt_cam, float* c2w_);
  void cam2base(float* pt_cam, float* pt_base, float* c2w_);
  Vertex VertexInterp(float isolevel, Vertex p1, Vertex p2, float valp1, float valp2);
}
extern "C" {
void ref_frame2cam(int px, int py, float z, float* K, float* out3) { int p[2] = {px, py}; ark::frame2cam(p, z, out3, K); }
void ref_cam2frame(float* cam3, float* K, int* out2) { ark::cam2frame(cam3, out2, K); }
void ref_base2cam(float* base3, float* c2w, float* out3) { ark::base2cam(base3, out3, c2w); }
void ref_cam2base(float* cam3, float* c2w, float* out3) { ark::cam2base(cam3, out3, c2w); }
void ref_vertex_interp(const float* p1, const float* p2, float v1, float v2, float* out3) {
  Vertex a(p1[0], p1[1], p1[2]), b(p2[0], p2[1], p2[2]); a.r = a.g = a.b = 0; b.r = b.g = b.b = 0;
  Vertex r = ark::VertexInterp(0.f, a, b, v1, v2); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
unsigned long long ref_block_hash(int x, int y, int z) { return (unsigned long long)BlockHasher()(make_int3(x, y, z)); }
}
