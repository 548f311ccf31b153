/* vh_oracle.c — CPU restatement of the reference's voxel-hashing TSDF hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this. The product (libvhsdf.so)
 * never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED. This restatement is checked bit-for-bit against the reference's own
 * src/tsdf.cu executed on the CPU through oracle/build_ref.sh (sequential CUDA emulation,
 * oracle/_ref/libref_emu_vpb{5,8}.so) — visible block lists per frame, every stored voxel's
 * sdf/weight/rgb, and the ordered triangle soup — by tests/test_oracle_vs_reference.py, and
 * against fixtures generated from that emulated reference (tests/golden/, made by
 * tests/golden/make_golden.py). The reference ships no golden vectors of its own.
 *
 * What it restates (all citations are /root/reference/...):
 *   A.1 candidate chunks    src/tsdf.cu:277-457 (streamInCPU2GPU), :154-187, :189-206
 *   A.2 visible block set   src/tsdf.cu:2088-2238 (HashAssignKernel), :2013-2064 (frustum test)
 *   A.3 integrate           src/tsdf.cu:599-751  (IntegrateHashKernel), :67-116 (camera math)
 *   A.4 marching cubes      src/tsdf.cu:884-1110 (marchingCubeHashKernel), :1640-1660 (VertexInterp)
 *   A.5 persistence         src/tsdf.cu:469-596  (streamOutGPU2CPU)
 *   A.6 mesh assembly       src/tsdf.cu:1760-1888 (tsdf2mesh)
 * Every quirk of SURVEY.md A.7 (Q1 transposed depth gate, Q2 weight-after-increment update,
 * Q3 unsigned MC thread mapping, Q4 no weight test, Q5 degenerate test, ...) is kept literally.
 *
 * Arithmetic: IEEE-754 binary32, no FMA contraction (build with -ffp-contract=off), expression
 * order as in the reference source. Unlike the reference every compile-time constant is a runtime
 * parameter (voxels per block, DDA stride, ray step cap, chunk radius, world extent).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vh_mc_tables.h"

typedef struct vo_params {
  int width, height;
  float fx, fy, cx, cy;
  float min_depth;        /* reference: 0.1, tsdf.cu:1318 */
  float max_depth;
  float vox_size;
  float trunc_margin;
  int voxels_per_block;   /* VOXEL_PER_BLOCK, tsdf.cuh:40 (reference 5) */
  int blocks_per_chunk;   /* BLOCK_PER_CHUNK, tsdf.cuh:41 */
  int dda_stride;         /* DDA_STEP, tsdf.cu:13 */
  int max_ray_steps;      /* maxLoopIterCount, tsdf.cu:2156 */
  float chunk_radius;     /* CHUNK_RADIUS, tsdf.cuh:44 */
  int max_chunk_num;      /* MAX_CHUNK_NUM, tsdf.cuh:43; 0 = unbounded world */
  int use_color;          /* 0: rgb input ignored, colours stay 0 */
  int run_mc;             /* 1: per-frame working-set marching cubes as processFrame does */
  int num_threads;        /* 0 = all OpenMP threads */
} vo_params;

typedef struct { int x, y, z; } vo_int3;
typedef struct { float x, y, z; } vo_float3;

typedef struct vo_tri {
  float p[9];
  unsigned char c[9];
  int slot;               /* tid*5 + k: the reference's dense slot inside the block, tsdf.cu:1042 */
} vo_tri;

typedef struct vo_block {
  vo_int3 key;
  float* sdf;
  float* w;
  unsigned char* rgb;     /* 3 per voxel */
  int stamp;              /* frame number (1-based) in which the block was last visible */
  int ntri, captri;
  vo_tri* tris;           /* triangles of the last frame that saw the block (A.5) */
} vo_block;

typedef struct vo_oracle {
  vo_params P;
  int nvox;               /* vpb^3 */
  float block_size, chunk_size;
  /* sparse map */
  vo_block* blocks; size_t nblocks, capblocks;
  int64_t* ht_key; int32_t* ht_val; size_t ht_cap; /* open addressing on packed keys */
  /* per frame */
  int frame;              /* frames processed so far */
  int* visible; int nvisible, capvisible;   /* block indices in discovery order */
  long long last_updates, last_tris, last_streamed_blocks;
  double t_alloc, t_integrate, t_mc;
  /* frame constants (A.1) */
  vo_float3 fc; vo_int3 cstart, cend; float chunk_test_radius;
  float c2w[16];
  signed char tri[256][16]; unsigned char ntri_tab[256]; unsigned short edge_mask[256];
} vo_oracle;

/* ------------------------------------------------------------------------------------------- */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static inline int64_t pack_key(int x, int y, int z) {
  return (int64_t)((((uint64_t)(uint32_t)(x + (1 << 20)) & 0x1FFFFF) << 42) | (((uint64_t)(uint32_t)(y + (1 << 20)) & 0x1FFFFF) << 21) |
                   ((uint64_t)(uint32_t)(z + (1 << 20)) & 0x1FFFFF));
}
static inline uint64_t mix64(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }

static void ht_rebuild(vo_oracle* o, size_t cap) {
  free(o->ht_key); free(o->ht_val);
  o->ht_cap = cap;
  o->ht_key = (int64_t*)malloc(cap * sizeof(int64_t));
  o->ht_val = (int32_t*)malloc(cap * sizeof(int32_t));
  for (size_t i = 0; i < cap; i++) o->ht_key[i] = -1;
  for (size_t b = 0; b < o->nblocks; b++) {
    int64_t k = pack_key(o->blocks[b].key.x, o->blocks[b].key.y, o->blocks[b].key.z);
    size_t h = mix64((uint64_t)k) & (cap - 1);
    while (o->ht_key[h] != -1) h = (h + 1) & (cap - 1);
    o->ht_key[h] = k; o->ht_val[h] = (int32_t)b;
  }
}
static int find_block(const vo_oracle* o, int x, int y, int z) {
  int64_t k = pack_key(x, y, z);
  size_t h = mix64((uint64_t)k) & (o->ht_cap - 1);
  while (o->ht_key[h] != -1) {
    if (o->ht_key[h] == k) return o->ht_val[h];
    h = (h + 1) & (o->ht_cap - 1);
  }
  return -1;
}
static int get_or_create_block(vo_oracle* o, int x, int y, int z) {
  int b = find_block(o, x, y, z);
  if (b >= 0) return b;
  if (o->nblocks == o->capblocks) {
    o->capblocks = o->capblocks ? o->capblocks * 2 : 4096;
    o->blocks = (vo_block*)realloc(o->blocks, o->capblocks * sizeof(vo_block));
  }
  if ((o->nblocks + 1) * 2 > o->ht_cap) ht_rebuild(o, o->ht_cap * 2);
  vo_block* nb = &o->blocks[o->nblocks];
  memset(nb, 0, sizeof(*nb));
  nb->key.x = x; nb->key.y = y; nb->key.z = z;
  /* fresh blocks are all-zero voxels: Voxel() ctor, tsdf.cuh:126-128 */
  nb->sdf = (float*)calloc((size_t)o->nvox, sizeof(float));
  nb->w = (float*)calloc((size_t)o->nvox, sizeof(float));
  nb->rgb = (unsigned char*)calloc((size_t)o->nvox * 3, 1);
  int64_t k = pack_key(x, y, z);
  size_t h = mix64((uint64_t)k) & (o->ht_cap - 1);
  while (o->ht_key[h] != -1) h = (h + 1) & (o->ht_cap - 1);
  o->ht_key[h] = k; o->ht_val[h] = (int32_t)o->nblocks;
  return (int)o->nblocks++;
}

/* ---- camera math, tsdf.cu:67-116 ----------------------------------------------------------- */
static inline void frame2cam(int px, int py, float z, const vo_params* P, float cam[3]) {
  cam[2] = z;
  cam[0] = ((float)px - P->cx) * z / P->fx;            /* tsdf.cu:71 */
  cam[1] = ((float)py - P->cy) * z / P->fy;            /* tsdf.cu:72 */
}
static inline void cam2base(const float cam[3], const float* c2w, float base[3]) {   /* tsdf.cu:96-101 */
  base[0] = cam[0] * c2w[0] + cam[1] * c2w[1] + cam[2] * c2w[2] + c2w[3];
  base[1] = cam[0] * c2w[4] + cam[1] * c2w[5] + cam[2] * c2w[6] + c2w[7];
  base[2] = cam[0] * c2w[8] + cam[1] * c2w[9] + cam[2] * c2w[10] + c2w[11];
}
static inline void base2cam(const float base[3], const float* c2w, float cam[3]) {   /* tsdf.cu:82-93 */
  float t0 = base[0] - c2w[3], t1 = base[1] - c2w[7], t2 = base[2] - c2w[11];
  cam[0] = c2w[0] * t0 + c2w[4] * t1 + c2w[8] * t2;
  cam[1] = c2w[1] * t0 + c2w[5] * t1 + c2w[9] * t2;
  cam[2] = c2w[2] * t0 + c2w[6] * t1 + c2w[10] * t2;
}
static inline vo_float3 frame2base(int px, int py, float z, const vo_params* P, const float* c2w) {
  float cam[3], base[3];
  frame2cam(px, py, z, P, cam);
  cam2base(cam, c2w, base);
  vo_float3 r = {base[0], base[1], base[2]};
  return r;
}

/* ---- A.1 candidate chunks ------------------------------------------------------------------ */
static void frame_setup(vo_oracle* o, const float* c2w) {
  const vo_params* P = &o->P;
  memcpy(o->c2w, c2w, sizeof(o->c2w));
  o->fc = frame2base(P->width / 2, P->height / 2, P->max_depth / 2, P, c2w);          /* tsdf.cu:154-161 */
  vo_int3 cc;                                                                          /* tsdf.cu:197-206 */
  cc.x = (int)floorf(o->fc.x / o->chunk_size);
  cc.y = (int)floorf(o->fc.y / o->chunk_size);
  cc.z = (int)floorf(o->fc.z / o->chunk_size);
  int rng = (int)ceil((double)P->chunk_radius / (double)o->chunk_size);               /* tsdf.cu:300 */
  int lo = P->max_chunk_num ? -P->max_chunk_num / 2 : INT32_MIN / 2;
  int hi = P->max_chunk_num ? P->max_chunk_num / 2 - 1 : INT32_MAX / 2;
  o->cstart.x = cc.x - rng > lo ? cc.x - rng : lo; o->cend.x = cc.x + rng < hi ? cc.x + rng : hi;   /* tsdf.cu:304-312 */
  o->cstart.y = cc.y - rng > lo ? cc.y - rng : lo; o->cend.y = cc.y + rng < hi ? cc.y + rng : hi;
  o->cstart.z = cc.z - rng > lo ? cc.z - rng : lo; o->cend.z = cc.z + rng < hi ? cc.z + rng : hi;
  /* float chunkRadius = 0.5f*CHUNK_RADIUS*sqrt(3.0f)*1.1 : double product stored to float, tsdf.cu:172 */
  o->chunk_test_radius = (float)((double)0.5f * (double)P->chunk_radius * (double)sqrtf(3.0f) * 1.1);
}
static inline int chunk_is_candidate(const vo_oracle* o, int x, int y, int z) {
  if (x < o->cstart.x || x > o->cend.x || y < o->cstart.y || y > o->cend.y || z < o->cstart.z || z > o->cend.z) return 0;
  /* tsdf.cu:166-187: centre = ((float)x + 0.5) * chunk_size in double, narrowed to float */
  float ccx = (float)(((double)(float)x + 0.5) * (double)o->chunk_size);
  float ccy = (float)(((double)(float)y + 0.5) * (double)o->chunk_size);
  float ccz = (float)(((double)(float)z + 0.5) * (double)o->chunk_size);
  float vx = o->fc.x - ccx, vy = o->fc.y - ccy, vz = o->fc.z - ccz;
  float l = sqrtf(vx * vx + vz * vz + vy * vy);                                        /* tsdf.cu:174 (x, z, y order) */
  return l <= fabsf(o->chunk_test_radius);
}
static inline int block2chunk1(int b, int bpc) { return (int)floorf((float)b / (float)bpc); }   /* tsdf.cu:256-260 */
static inline int block_is_candidate(const vo_oracle* o, int bx, int by, int bz) {
  int bpc = o->P.blocks_per_chunk;
  return chunk_is_candidate(o, block2chunk1(bx, bpc), block2chunk1(by, bpc), block2chunk1(bz, bpc));
}
static long long count_streamed_blocks(const vo_oracle* o) {
  long long n = 0; int bpc = o->P.blocks_per_chunk;
  for (int x = o->cstart.x; x <= o->cend.x; x++) for (int y = o->cstart.y; y <= o->cend.y; y++) for (int z = o->cstart.z; z <= o->cend.z; z++)
    if (chunk_is_candidate(o, x, y, z)) n += (long long)bpc * bpc * bpc;
  return n;
}

/* ---- A.2.1 frustum test, tsdf.cu:2013-2064 ------------------------------------------------- */
static inline int block_in_frustum(const vo_oracle* o, float px, float py, float pz) {
  const vo_params* P = &o->P; const float* c2w = o->c2w;
  float t0 = px - c2w[3], t1 = py - c2w[7], t2 = pz - c2w[11];
  float cx_ = c2w[0] * t0 + c2w[4] * t1 + c2w[8] * t2;
  float cy_ = c2w[1] * t0 + c2w[5] * t1 + c2w[9] * t2;
  float cz_ = c2w[2] * t0 + c2w[6] * t1 + c2w[10] * t2;
  float u = cx_ * P->fx / cz_ + P->cx;                                                 /* tsdf.cu:2016-2017 */
  float v = cy_ * P->fy / cz_ + P->cy;
  float wm1 = (float)P->width - 1.0f, hm1 = (float)P->height - 1.0f;
  float ix = (2.0f * u - wm1) / wm1;                                                   /* tsdf.cu:2031 */
  float iy = (hm1 - 2.0f * v) / hm1;                                                   /* tsdf.cu:2032 */
  float iz = (cz_ - P->min_depth) / (P->max_depth - P->min_depth);                     /* tsdf.cu:2022 */
  const float k = (float)0.95;                                                         /* pProj *= 0.95, tsdf.cu:2062 */
  ix *= k; iy *= k; iz *= k;
  return !(ix < -1.0f || ix > 1.0f || iy < -1.0f || iy > 1.0f || iz < 0.0f || iz > 1.0f);
}

static inline float sgnf(float v) { return (float)((0.0f < v) - (v < 0.0f)); }         /* cutil_math.h:31-33 */
static inline float fminf_ref(float a, float b) { return fminf(a, b); }

/* ---- A.2 visible set ----------------------------------------------------------------------- */
static void mark_visible(vo_oracle* o, int bx, int by, int bz) {
  int b = get_or_create_block(o, bx, by, bz);
  if (o->blocks[b].stamp == o->frame) return;
  o->blocks[b].stamp = o->frame;
  if (o->nvisible == o->capvisible) {
    o->capvisible = o->capvisible ? o->capvisible * 2 : 8192;
    o->visible = (int*)realloc(o->visible, (size_t)o->capvisible * sizeof(int));
  }
  o->visible[o->nvisible++] = b;
}

static void cast_ray(vo_oracle* o, const float* depth, unsigned x, unsigned y) {
  const vo_params* P = &o->P; const float* c2w = o->c2w;
  const float bs = o->block_size, vs = P->vox_size;
  /* Q1: depth[x*width + y] with x = column (tsdf.cu:2114); out of range reads are 0 (zero-padded buffer) */
  size_t idx = (size_t)x * (size_t)P->width + (size_t)y;
  float d = idx < (size_t)P->width * (size_t)P->height ? depth[idx] : 0.0f;
  if (d == 0.0f || d == -INFINITY) return;                                             /* tsdf.cu:2116 */
  if (d >= P->max_depth) return;                                                       /* tsdf.cu:2119 */
  float t = P->trunc_margin;
  float mind = fminf_ref(P->max_depth, d - t), maxd = fminf_ref(P->max_depth, d + t);  /* tsdf.cu:2122-2126 */
  if (mind >= maxd) return;
  vo_float3 rmin = frame2base((int)x, (int)y, P->min_depth, P, c2w);                   /* tsdf.cu:2129-2130 */
  vo_float3 rmax = frame2base((int)x, (int)y, P->max_depth, P, c2w);
  float vx = rmax.x - rmin.x, vy = rmax.y - rmin.y, vz = rmax.z - rmin.z;
  float inv = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);                               /* rsqrtf host form, cutil_math.h:81-84,1207-1211 */
  float dir[3] = {vx * inv, vy * inv, vz * inv};
  float rm[3] = {rmin.x, rmin.y, rmin.z}, rM[3] = {rmax.x, rmax.y, rmax.z};
  int cur[3], bound[3]; float step[3], tmax[3], tdel[3];
  for (int a = 0; a < 3; a++) {
    cur[a] = (int)floorf(rm[a] / bs);                                                   /* tsdf.cu:2136, :48-54 */
    int end = (int)floorf(rM[a] / bs);
    step[a] = sgnf(dir[a]);                                                             /* tsdf.cu:2145 */
    float cl = fmaxf(0.0f, fminf(step[a], 1.0f));                                       /* clamp(step,0,1), cutil_math.h:1050 */
    float boundary = (float)(cur[a] + (int)cl) * bs - 0.5f * vs;                        /* tsdf.cu:2146 */
    tmax[a] = (boundary - rm[a]) / dir[a];                                              /* tsdf.cu:2147 */
    tdel[a] = (step[a] * vs * (float)P->voxels_per_block) / dir[a];                     /* tsdf.cu:2148 */
    bound[a] = (int)((float)end + step[a]);                                             /* tsdf.cu:2149 */
    if (dir[a] == 0.0f || boundary - rm[a] == 0.0f) { tmax[a] = INFINITY; tdel[a] = INFINITY; }   /* tsdf.cu:2151-2153 */
  }
  for (int iter = 0; iter < P->max_ray_steps; iter++) {                                 /* tsdf.cu:2158 */
    if (block_is_candidate(o, cur[0], cur[1], cur[2])) {                                /* find() in the streamed-in table, :2164 */
      if (block_in_frustum(o, (float)cur[0] * bs, (float)cur[1] * bs, (float)cur[2] * bs))   /* :2165, block2world :40-46 */
        mark_visible(o, cur[0], cur[1], cur[2]);
    }
    int a;                                                                              /* tsdf.cu:2217-2233 */
    if (tmax[0] < tmax[1] && tmax[0] < tmax[2]) a = 0;
    else if (tmax[2] < tmax[1]) a = 2;
    else a = 1;
    cur[a] = (int)((float)cur[a] + step[a]);
    if (cur[a] == bound[a]) return;
    tmax[a] += tdel[a];
  }
}

static void allocate_visible(vo_oracle* o, const float* depth) {
  const vo_params* P = &o->P;
  o->nvisible = 0;
  /* launch shape of HashAssign, tsdf.cu:2263-2264: threads of 8x8 CUDA blocks, pixel = thread * DDA_STEP.
   * Iterated in the emulation's sequential order (block y, block x, thread y, thread x). */
  const int T = 8;
  int gx = (P->width / P->dda_stride + T - 1) / T, gy = (P->height / P->dda_stride + T - 1) / T;
  for (int by = 0; by < gy; by++) for (int bx = 0; bx < gx; bx++)
    for (int ty = 0; ty < T; ty++) for (int tx = 0; tx < T; tx++) {
      unsigned x = (unsigned)(bx * T + tx) * (unsigned)P->dda_stride;
      unsigned y = (unsigned)(by * T + ty) * (unsigned)P->dda_stride;
      if (x < (unsigned)P->width && y < (unsigned)P->height) cast_ray(o, depth, x, y);  /* tsdf.cu:2110 */
    }
}

/* ---- A.3 integrate ------------------------------------------------------------------------- */
static long long integrate_block(const vo_oracle* o, vo_block* b, const float* depth, const unsigned char* rgb) {
  const vo_params* P = &o->P; const float* c2w = o->c2w;
  const int V = P->voxels_per_block, V2 = V * V;
  long long updates = 0;
  for (int tid = 0; tid < o->nvox; tid++) {
    int z = tid % V, y = ((tid - z) % V2) / V, x = tid / V2;                            /* tsdf.cu:614-617 */
    float pb[3] = {(float)(b->key.x * V + x) * P->vox_size, (float)(b->key.y * V + y) * P->vox_size,
                   (float)(b->key.z * V + z) * P->vox_size};                            /* tsdf.cu:621-623, :24-30 */
    float cam[3];
    base2cam(pb, c2w, cam);                                                             /* tsdf.cu:684 */
    float fu = roundf(P->fx * (cam[0] / cam[2]) + P->cx);                               /* tsdf.cu:76-79 */
    float fv = roundf(P->fy * (cam[1] / cam[2]) + P->cy);
    if (cam[2] <= 0) continue;                                                          /* tsdf.cu:706 */
    /* the reference converts to int first; out-of-range / NaN conversions land outside the image
     * on both x86 (INT_MIN) and CUDA (saturation), so the float-domain test is equivalent */
    if (!(fu >= 0.0f && fu < (float)P->width && fv >= 0.0f && fv < (float)P->height)) continue;   /* tsdf.cu:710 */
    int u = (int)fu, v = (int)fv;
    float dv = depth[v * P->width + u];                                                 /* tsdf.cu:713 */
    if (dv <= 0 || dv > P->max_depth) continue;                                         /* tsdf.cu:715 */
    float diff = dv - cam[2];
    if (diff <= -P->trunc_margin) continue;                                             /* tsdf.cu:720 */
    int img = v * P->width + u;
    float dist = fminf(1.0f, diff / P->trunc_margin);                                   /* tsdf.cu:738 */
    float w_old = b->w[tid], w_new = w_old + 1.0f;
    b->w[tid] = w_new;
    b->sdf[tid] = (b->sdf[tid] * w_new + dist) / w_new;                                 /* Q2, tsdf.cu:741-742 */
    if (P->use_color && rgb) {
      for (int k = 0; k < 3; k++)                                                       /* tsdf.cu:743-745 */
        b->rgb[3 * tid + k] = (unsigned char)(((float)b->rgb[3 * tid + k] * w_old + (float)rgb[3 * img + k]) / w_new);
    }
    updates++;
  }
  return updates;
}

/* ---- A.4 marching cubes -------------------------------------------------------------------- */
typedef struct { float x, y, z; unsigned char r, g, b; } vo_vertex;

static vo_vertex vertex_interp(vo_vertex p1, vo_vertex p2, float v1, float v2) {        /* tsdf.cu:1640-1660, isolevel 0 */
  if (fabs(0.0f - v1) < 0.00001) return p1;
  if (fabs(0.0f - v2) < 0.00001) return p2;
  if (fabs(v1 - v2) < 0.00001) return p1;
  float mu = (0.0f - v1) / (v2 - v1);
  vo_vertex p;
  p.x = p1.x + mu * (p2.x - p1.x);
  p.y = p1.y + mu * (p2.y - p1.y);
  p.z = p1.z + mu * (p2.z - p1.z);
  p.r = (unsigned char)((float)p1.r + mu * (float)((int)p2.r - (int)p1.r));
  p.g = (unsigned char)((float)p1.g + mu * (float)((int)p2.g - (int)p1.g));
  p.b = (unsigned char)((float)p1.b + mu * (float)((int)p2.b - (int)p1.b));
  return p;
}
static inline int vtx_eq(const vo_vertex* a, const vo_vertex* b) { return a->x == b->x && a->y == b->y && a->z == b->z; }

static void mc_block(const vo_oracle* o, vo_block* b) {
  const int V = o->P.voxels_per_block; const unsigned V2 = (unsigned)(V * V);
  b->ntri = 0;
  for (unsigned tid = 0; tid < (unsigned)o->nvox; tid++) {
    /* Q3: unsigned arithmetic exactly as written, tsdf.cu:903-906 */
    int z = (int)(tid % (unsigned)V + (unsigned)(b->key.z * V));
    int y = (int)(((tid - (unsigned)z) % V2) / (unsigned)V + (unsigned)(b->key.y * V));
    int x = (int)(tid / V2 + (unsigned)(b->key.x * V));
    vo_vertex gp[8]; float val[8]; int ok = 1;
    for (int k = 0; k < 8 && ok; k++) {
      int cx = x + VH_MC_CORNER_OFFSET[k][0], cy = y + VH_MC_CORNER_OFFSET[k][1], cz = z + VH_MC_CORNER_OFFSET[k][2];
      int nx = (int)floorf((float)cx / (float)V), ny = (int)floorf((float)cy / (float)V), nz = (int)floorf((float)cz / (float)V);   /* :56-58 */
      int lin = ((cx - nx * V) * V + (cy - ny * V)) * V + (cz - nz * V);               /* tsdf.cu:60-64 */
      int nb = find_block(o, nx, ny, nz);
      if (nb < 0 || o->blocks[nb].stamp != o->frame) { ok = 0; break; }                 /* not in this frame's working set, :930,:966 */
      const vo_block* B = &o->blocks[nb];
      gp[k].x = (float)cx; gp[k].y = (float)cy; gp[k].z = (float)cz;
      gp[k].r = B->rgb[3 * lin]; gp[k].g = B->rgb[3 * lin + 1]; gp[k].b = B->rgb[3 * lin + 2];
      val[k] = B->sdf[lin];                                                             /* Q4: no weight test */
    }
    if (!ok) continue;
    int cube = 0;
    for (int k = 0; k < 8; k++) if (val[k] < 0) cube |= 1 << k;                          /* tsdf.cu:978-986 */
    unsigned em = o->edge_mask[cube];
    if (em == 0) continue;
    vo_vertex vl[12];
    for (int e = 0; e < 12; e++)
      if (em & (1u << e)) { int a = VH_MC_EDGE_CORNERS[e][0], c = VH_MC_EDGE_CORNERS[e][1]; vl[e] = vertex_interp(gp[a], gp[c], val[a], val[c]); }
    int count = 0;
    for (int ti = 0; o->tri[cube][ti] != -1; ti += 3, count++) {                        /* tsdf.cu:1044-1061 */
      const vo_vertex* p0 = &vl[o->tri[cube][ti]], *p1 = &vl[o->tri[cube][ti + 1]], *p2 = &vl[o->tri[cube][ti + 2]];
      if (vtx_eq(p0, p1) || vtx_eq(p1, p2) || vtx_eq(p0, p1)) continue;                 /* Q5 */
      if (b->ntri == b->captri) { b->captri = b->captri ? b->captri * 2 : 16; b->tris = (vo_tri*)realloc(b->tris, (size_t)b->captri * sizeof(vo_tri)); }
      vo_tri* t = &b->tris[b->ntri++];
      const vo_vertex* pp[3] = {p0, p1, p2};
      for (int j = 0; j < 3; j++) {
        t->p[3 * j] = pp[j]->x; t->p[3 * j + 1] = pp[j]->y; t->p[3 * j + 2] = pp[j]->z;
        t->c[3 * j] = pp[j]->r; t->c[3 * j + 1] = pp[j]->g; t->c[3 * j + 2] = pp[j]->b;
      }
      t->slot = (int)tid * 5 + count;
    }
  }
}

/* ---- public API ---------------------------------------------------------------------------- */
void vo_default_params(vo_params* p) {
  memset(p, 0, sizeof(*p));
  p->width = 640; p->height = 480; p->fx = p->fy = 577.0f; p->cx = 320.0f; p->cy = 240.0f;
  p->min_depth = 0.1f; p->max_depth = 10.0f; p->vox_size = 0.01f; p->trunc_margin = 0.05f;
  p->voxels_per_block = 8; p->blocks_per_chunk = 8; p->dda_stride = 10; p->max_ray_steps = 100;
  p->chunk_radius = 4.0f; p->max_chunk_num = 128; p->use_color = 1; p->run_mc = 1; p->num_threads = 0;
}

vo_oracle* vo_create(const vo_params* p) {
  vo_oracle* o = (vo_oracle*)calloc(1, sizeof(vo_oracle));
  o->P = *p;
  o->nvox = p->voxels_per_block * p->voxels_per_block * p->voxels_per_block;
  o->block_size = (float)p->voxels_per_block * p->vox_size;                                          /* tsdf.cu:1326 */
  o->chunk_size = (float)(p->blocks_per_chunk * p->voxels_per_block) * p->vox_size;                  /* tsdf.cu:1271 */
  vh_mc_expand_tables(o->tri, o->ntri_tab, o->edge_mask);
  ht_rebuild(o, 1 << 16);
  return o;
}

void vo_destroy(vo_oracle* o) {
  if (!o) return;
  for (size_t i = 0; i < o->nblocks; i++) { free(o->blocks[i].sdf); free(o->blocks[i].w); free(o->blocks[i].rgb); free(o->blocks[i].tris); }
  free(o->blocks); free(o->ht_key); free(o->ht_val); free(o->visible); free(o);
}

/* One processFrame (tsdf.cu:1485-1598): candidates, visible set, integrate, working-set MC, persist. */
void vo_process_frame(vo_oracle* o, const float* depth, const unsigned char* rgb, const float* c2w) {
  o->frame++;
  double t0 = now_s();
  frame_setup(o, c2w);
  allocate_visible(o, depth);
  double t1 = now_s();
  long long updates = 0;
  int nt = o->P.num_threads;
#ifdef _OPENMP
  if (nt <= 0) nt = omp_get_max_threads();
#else
  nt = 1;
#endif
  (void)nt;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : updates) num_threads(nt)
  for (int i = 0; i < o->nvisible; i++) updates += integrate_block(o, &o->blocks[o->visible[i]], depth, rgb);
  double t2 = now_s();
  long long tris = 0;
  if (o->P.run_mc) {
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tris) num_threads(nt)
    for (int i = 0; i < o->nvisible; i++) { vo_block* b = &o->blocks[o->visible[i]]; mc_block(o, b); tris += b->ntri; }
  }
  double t3 = now_s();
  o->last_updates = updates; o->last_tris = tris;
  o->t_alloc = t1 - t0; o->t_integrate = t2 - t1; o->t_mc = t3 - t2;
}

/* stage entry points for unit tests and for driving a GPU stage with oracle inputs */
void vo_begin_frame(vo_oracle* o, const float* c2w) { o->frame++; frame_setup(o, c2w); o->nvisible = 0; }
void vo_stage_allocate(vo_oracle* o, const float* depth) { allocate_visible(o, depth); }
long long vo_stage_integrate(vo_oracle* o, const float* depth, const unsigned char* rgb) {
  long long u = 0;
  for (int i = 0; i < o->nvisible; i++) u += integrate_block(o, &o->blocks[o->visible[i]], depth, rgb);
  o->last_updates = u; return u;
}
long long vo_stage_mc(vo_oracle* o) {
  long long t = 0;
  for (int i = 0; i < o->nvisible; i++) { mc_block(o, &o->blocks[o->visible[i]]); t += o->blocks[o->visible[i]].ntri; }
  o->last_tris = t; return t;
}

/* Full-map extraction (no reference counterpart: the reference only ever meshes a frame's working set). Every allocated
 * block is meshed with the working-set rule applied to the whole map: a corner counts if its block is allocated at all.
 * Implemented as "a frame whose working set is every block"; it replaces the per-block meshes kept so far, so call it last. */
long long vo_full_map_mc(vo_oracle* o) {
  o->frame++;
  o->nvisible = 0;
  for (size_t b = 0; b < o->nblocks; b++) mark_visible(o, o->blocks[b].key.x, o->blocks[b].key.y, o->blocks[b].key.z);
  return vo_stage_mc(o);
}

int vo_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
int vo_num_visible(const vo_oracle* o) { return o->nvisible; }
long long vo_last_updates(const vo_oracle* o) { return o->last_updates; }
long long vo_last_triangles(const vo_oracle* o) { return o->last_tris; }
long long vo_streamed_blocks(const vo_oracle* o) { return count_streamed_blocks(o); }
void vo_last_times(const vo_oracle* o, double* t3) { t3[0] = o->t_alloc; t3[1] = o->t_integrate; t3[2] = o->t_mc; }
int vo_visible_keys(const vo_oracle* o, int* out_xyz, int cap) {
  for (int i = 0; i < o->nvisible && i < cap; i++) { const vo_int3* k = &o->blocks[o->visible[i]].key; out_xyz[3 * i] = k->x; out_xyz[3 * i + 1] = k->y; out_xyz[3 * i + 2] = k->z; }
  return o->nvisible;
}
/* blocks that have ever been visible (= the persistent allocated set of the new engine) */
long long vo_num_blocks(const vo_oracle* o) { return (long long)o->nblocks; }
long long vo_all_keys(const vo_oracle* o, int* out_xyz, long long cap) {
  for (long long i = 0; i < (long long)o->nblocks && i < cap; i++) { out_xyz[3 * i] = o->blocks[i].key.x; out_xyz[3 * i + 1] = o->blocks[i].key.y; out_xyz[3 * i + 2] = o->blocks[i].key.z; }
  return (long long)o->nblocks;
}
int vo_get_block(const vo_oracle* o, int bx, int by, int bz, float* sdf, float* w, unsigned char* rgb) {
  int b = find_block(o, bx, by, bz);
  if (b < 0) return 0;
  if (sdf) memcpy(sdf, o->blocks[b].sdf, (size_t)o->nvox * sizeof(float));
  if (w) memcpy(w, o->blocks[b].w, (size_t)o->nvox * sizeof(float));
  if (rgb) memcpy(rgb, o->blocks[b].rgb, (size_t)o->nvox * 3);
  return 1;
}
/* bulk form: keys[n] -> sdf[n*nvox], w[n*nvox], rgb[n*nvox*3]; found[n] */
void vo_get_blocks(const vo_oracle* o, const int* keys_xyz, long long n, float* sdf, float* w, unsigned char* rgb, unsigned char* found) {
  for (long long i = 0; i < n; i++) {
    found[i] = (unsigned char)vo_get_block(o, keys_xyz[3 * i], keys_xyz[3 * i + 1], keys_xyz[3 * i + 2], sdf ? sdf + i * o->nvox : 0,
                                           w ? w + i * o->nvox : 0, rgb ? rgb + i * o->nvox * 3 : 0);
  }
}
void vo_voxel_checksum(const vo_oracle* o, double* sum_sdf, double* sum_w, long long* n_observed, long long* n_negative) {
  double ss = 0, sw = 0; long long no = 0, nn = 0;
  for (size_t b = 0; b < o->nblocks; b++) for (int v = 0; v < o->nvox; v++) {
    ss += o->blocks[b].sdf[v]; sw += o->blocks[b].w[v]; no += o->blocks[b].w[v] > 0; nn += o->blocks[b].sdf[v] < 0;
  }
  *sum_sdf = ss; *sum_w = sw; *n_observed = no; *n_negative = nn;
}

/* A.6 order: chunks x,y,z ascending, block linear index inside the chunk, slot ascending (tsdf.cu:1786-1806) */
static const vo_oracle* g_sort_o;
static int cmp_block_mesh_order(const void* a, const void* b) {
  const vo_block* A = &g_sort_o->blocks[*(const int*)a]; const vo_block* B = &g_sort_o->blocks[*(const int*)b];
  int bpc = g_sort_o->P.blocks_per_chunk;
  int ca[3] = {block2chunk1(A->key.x, bpc), block2chunk1(A->key.y, bpc), block2chunk1(A->key.z, bpc)};
  int cb[3] = {block2chunk1(B->key.x, bpc), block2chunk1(B->key.y, bpc), block2chunk1(B->key.z, bpc)};
  for (int i = 0; i < 3; i++) if (ca[i] != cb[i]) return ca[i] < cb[i] ? -1 : 1;
  int ka[3] = {A->key.x, A->key.y, A->key.z}, kb[3] = {B->key.x, B->key.y, B->key.z};
  for (int i = 0; i < 3; i++) if (ka[i] != kb[i]) return ka[i] < kb[i] ? -1 : 1;   /* same chunk: (lx*bpc+ly)*bpc+lz order */
  return 0;
}
static int cmp_tri_slot(const void* a, const void* b) { return ((const vo_tri*)a)->slot - ((const vo_tri*)b)->slot; }

/* Final mesh as the ordered triangle soup tsdf2mesh walks (voxel-index units). Returns the count. */
long long vo_triangles(vo_oracle* o, float* out_xyz, unsigned char* out_rgb, long long cap) {
  int* order = (int*)malloc((o->nblocks ? o->nblocks : 1) * sizeof(int)); size_t m = 0;
  for (size_t b = 0; b < o->nblocks; b++) if (o->blocks[b].ntri > 0) order[m++] = (int)b;
  g_sort_o = o;
  qsort(order, m, sizeof(int), cmp_block_mesh_order);
  long long n = 0;
  for (size_t i = 0; i < m; i++) {
    vo_block* b = &o->blocks[order[i]];
    qsort(b->tris, (size_t)b->ntri, sizeof(vo_tri), cmp_tri_slot);
    for (int t = 0; t < b->ntri; t++, n++) {
      if (n >= cap) continue;
      if (out_xyz) memcpy(out_xyz + 9 * n, b->tris[t].p, sizeof(float) * 9);
      if (out_rgb) memcpy(out_rgb + 9 * n, b->tris[t].c, 9);
    }
  }
  free(order);
  return n;
}

/* per-block triangle counts of the last frame's working set, in visible order */
void vo_visible_tri_counts(const vo_oracle* o, int* out) { for (int i = 0; i < o->nvisible; i++) out[i] = o->blocks[o->visible[i]].ntri; }

/* known-answer hooks for the shared math */
void vo_frame2base(const vo_params* P, const float* c2w, int px, int py, float z, float* out3) { vo_float3 r = frame2base(px, py, z, P, c2w); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }
void vo_base2cam(const float* base3, const float* c2w, float* out3) { base2cam(base3, c2w, out3); }
void vo_cam2frame(const vo_params* P, const float* cam3, float* out2) {
  out2[0] = roundf(P->fx * (cam3[0] / cam3[2]) + P->cx); out2[1] = roundf(P->fy * (cam3[1] / cam3[2]) + P->cy);
}
void vo_vertex_interp(const float* p1, const float* p2, float v1, float v2, float* out3) {
  vo_vertex a = {p1[0], p1[1], p1[2], 0, 0, 0}, b = {p2[0], p2[1], p2[2], 0, 0, 0};
  vo_vertex r = vertex_interp(a, b, v1, v2); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
/* BlockHasher, tsdf.cuh:144-156: sign-extended 64-bit products xor-ed */
unsigned long long vo_block_hash(int x, int y, int z) {
  return ((unsigned long long)(long long)x * 73856093ULL) ^ ((unsigned long long)(long long)y * 19349669ULL) ^ ((unsigned long long)(long long)z * 83492791ULL);
}
int vo_block_is_candidate(vo_oracle* o, const float* c2w, int bx, int by, int bz) { frame_setup(o, c2w); return block_is_candidate(o, bx, by, bz); }
int vo_block_in_frustum(vo_oracle* o, const float* c2w, int bx, int by, int bz) {
  memcpy(o->c2w, c2w, sizeof(o->c2w));
  return block_in_frustum(o, (float)bx * o->block_size, (float)by * o->block_size, (float)bz * o->block_size);
}
