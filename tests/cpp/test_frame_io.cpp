// test_frame_io.cpp — CPU-only check of the frame I/O drop-in (include/SaveFrame.h, include/Utils.h): load frame 0 of a
// ScanNet-layout folder the way load_frames does (SaveFrame.cpp:154-218), dump what the engine would receive, and
// write the frame back under id 1 (frameWrite, SaveFrame.cpp:120-152). tests/test_cpp_host.py compares both with cv2.
#include <SaveFrame.h>

#include <cstdio>

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  ark::SaveFrame io(argv[1]);
  ark::RGBDFrame f = io.frameLoad(0);
  if (f.frameId != 0) { std::fprintf(stderr, "frame 0 not loaded\n"); return 1; }
  if (io.frameLoad(7).frameId != -1) { std::fprintf(stderr, "missing frame must report -1\n"); return 1; }
  const ark::Mat twc = f.mTcw.inv();      // PushFrame's inversion (PointCloudGenerator.cpp:140)
  FILE* o = std::fopen((std::string(argv[1]) + "dump.bin").c_str(), "wb");
  const int dims[4] = {f.imDepth.cols, f.imDepth.rows, f.imRGB.cols, f.imRGB.rows};
  std::fwrite(dims, sizeof(int), 4, o);
  std::fwrite(f.imDepth.datastart, sizeof(float), (size_t)dims[0] * dims[1], o);
  std::fwrite(f.imRGB.datastart, 1, (size_t)dims[2] * dims[3] * 3, o);
  std::fwrite(twc.datastart, sizeof(float), 16, o);
  std::fclose(o);
  f.frameId = 1;
  io.frameWrite(f);
  std::printf("FRAME_IO_OK\n");
  return 0;
}
