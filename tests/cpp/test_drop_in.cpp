// test_drop_in.cpp — the reference's caller-side code path against this repo's drop-in header.
// Mirrors what PointCloudGenerator::Reproject does with the engine (/root/reference/src/PointCloudGenerator.cpp:43-47,
// :127, :154-158): construct ark::GpuTsdfGenerator with the 15 reference ctor arguments, call processFrame with raw
// host pointers per frame, SavePLY at the end. Plain C++ (g++), links libvhsdf.so only.
//
// usage: test_drop_in <frames.bin> <out_dir>
//   frames.bin: int32 {W, H, n}; float32 {fx, fy, cx, cy, max_depth, vox_size, trunc}; then per frame
//               float32 depth[H*W], uint8 rgb[H*W*3], float32 c2w[16]       (written by tests/test_cpp_drop_in.py)
#include <tsdf.cuh>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s frames.bin out_dir\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror(argv[1]); return 2; }
  int hdr[3]; float cam[7];
  if (std::fread(hdr, sizeof(int), 3, f) != 3 || std::fread(cam, sizeof(float), 7, f) != 7) return 2;
  const int W = hdr[0], H = hdr[1], n = hdr[2];
  const std::string out = argv[2];

  vh_params& opt = ark::GpuTsdfGenerator::Options();       // run-time versions of the reference's macros
  opt.num_buckets = 1 << 16; opt.pool_blocks = 1 << 16; opt.tri_arena_bytes = 64ull << 20;

  try {
    ark::GpuTsdfGenerator gen(W, H, cam[0], cam[1], cam[2], cam[3], cam[4], 0.5f, 0.0f, -0.5f, cam[5], cam[6], 8, 8, 8);
    std::vector<float> depth((size_t)W * H);
    std::vector<unsigned char> rgb((size_t)W * H * 3);
    float c2w[16];
    for (int i = 0; i < n; i++) {
      if (std::fread(depth.data(), sizeof(float), depth.size(), f) != depth.size() || std::fread(rgb.data(), 1, rgb.size(), f) != rgb.size() ||
          std::fread(c2w, sizeof(float), 16, f) != 16) { std::fprintf(stderr, "short read in frame %d\n", i); return 2; }
      gen.processFrame(depth.data(), rgb.data(), c2w);
      const vh_stats s = gen.stats();
      std::printf("frame %d visible %u updates %llu triangles %llu\n", i, s.visible_blocks, (unsigned long long)s.voxel_updates,
                  (unsigned long long)s.triangles);
    }
    gen.SavePLY(out + "/model.ply");
    gen.SaveTSDF(out + "/tsdf.bin");
    std::printf("vertices %zu faces %zu\n", gen.getVertices()->size(), gen.getFaces()->size());
    ark::MarchingCubeParam* mp = gen.getMarchingCubeParam();
    std::printf("param vox_size %g trunc %g block_size %g min_depth %g\n", mp->vox_size, mp->trunc_margin, mp->block_size, mp->min_depth);
    // find_vertex finds what getVertices holds
    if (!gen.getVertices()->empty()) {
      std::vector<std::list<std::pair<ark::Vertex, int>>> dummy;
      const ark::Vertex v = (*gen.getVertices())[gen.getVertices()->size() / 2];
      const int idx = gen.find_vertex(v, make_uint3(1, 1, 1), 1.0f, dummy);
      if (idx < 0 || !((*gen.getVertices())[idx] == v)) { std::fprintf(stderr, "find_vertex failed\n"); return 1; }
    }
    gen.Shutdown();
    // error convention: a pool that is too small throws the reference's C string
    opt.pool_blocks = 32;
    bool threw = false;
    try {
      ark::GpuTsdfGenerator tiny(W, H, cam[0], cam[1], cam[2], cam[3], cam[4], 0, 0, 0, cam[5], cam[6], 8, 8, 8);
      std::fseek(f, 3 * sizeof(int) + 7 * sizeof(float), SEEK_SET);
      if (std::fread(depth.data(), sizeof(float), depth.size(), f) != depth.size() || std::fread(rgb.data(), 1, rgb.size(), f) != rgb.size() ||
          std::fread(c2w, sizeof(float), 16, f) != 16) return 2;
      tiny.processFrame(depth.data(), rgb.data(), c2w);
    } catch (const char* msg) {
      threw = std::string(msg) == "out of block memory";
    }
    if (!threw) { std::fprintf(stderr, "expected \"out of block memory\"\n"); return 1; }
    std::printf("DROP_IN_OK\n");
  } catch (const char* msg) {
    std::fprintf(stderr, "exception: %s (%s)\n", msg, vh_last_error());
    return 1;
  }
  std::fclose(f);
  return 0;
}
