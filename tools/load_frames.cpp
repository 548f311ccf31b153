// load_frames.cpp — headless form of the reference's executable (/root/reference/src/main.cpp): same arguments
// (frames folder with trailing '/', settings yaml), same frame loop as application_thread (main.cpp:236-283: frame ids
// 0, 20, 40, ... < 2025, stop after 5 missing frames, PushFrame each), then the 'p' key's action (main.cpp:321-327:
// RequestStop + SavePly("model.ply")). The GLUT window, camera keys and immediate-mode rendering are not part of the
// hot path and are left out. Timing uses a wall clock (the reference prints clock() CPU seconds and a hard-coded 101).
//
// usage: load_frames <frames_dir/> <settings.yaml> [--out model.ply] [--step 20] [--end 2025] [--poses poses.txt]
//                    [--buckets N] [--pool N] [--ray-steps N] [--no-mc]
#include <PointCloudGenerator.h>
#include <SaveFrame.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "Usage: " << argv[0] << " <frames_dir/> <settings.yaml> [--out model.ply] [--step 20] [--end 2025] [--poses file]" << std::endl;
    return -1;
  }
  std::string out = "model.ply", poses;
  int step = 20, end = 2025;
  vh_params& opt = ark::GpuTsdfGenerator::Options();
  for (int i = 3; i < argc; i++) {
    auto next = [&](const char* what) -> const char* { if (i + 1 >= argc) { std::cerr << what << " needs a value" << std::endl; exit(2); } return argv[++i]; };
    if (!std::strcmp(argv[i], "--out")) out = next("--out");
    else if (!std::strcmp(argv[i], "--step")) step = std::atoi(next("--step"));
    else if (!std::strcmp(argv[i], "--end")) end = std::atoi(next("--end"));
    else if (!std::strcmp(argv[i], "--poses")) poses = next("--poses");
    else if (!std::strcmp(argv[i], "--buckets")) opt.num_buckets = std::atoi(next("--buckets"));
    else if (!std::strcmp(argv[i], "--pool")) opt.pool_blocks = std::atoi(next("--pool"));
    else if (!std::strcmp(argv[i], "--ray-steps")) opt.max_ray_steps = std::atoi(next("--ray-steps"));
    else if (!std::strcmp(argv[i], "--no-mc")) opt.mc_per_frame = 0;
    else { std::cerr << "unknown option " << argv[i] << std::endl; return 2; }
  }
  try {
    ark::PointCloudGenerator pointCloudGenerator(argv[2]);
    ark::SaveFrame saveFrame(argv[1]);
    std::ofstream pose_log;
    if (!poses.empty()) { pose_log.open(poses); pose_log.precision(9); }

    pointCloudGenerator.Start();
    int tframe = 0, empty = 0, frames = 0;
    const auto t1 = std::chrono::steady_clock::now();
    while (tframe < end) {
      if (empty == 5) break;
      ark::RGBDFrame frame = saveFrame.frameLoad(tframe);
      tframe += step;
      if (frame.frameId == -1) { empty++; continue; }
      pointCloudGenerator.PushFrame(frame);
      frames++;
      if (pose_log.is_open()) {
        pose_log << frame.frameId;
        for (int k = 0; k < 16; k++) pose_log << " " << pointCloudGenerator.lastPose()[k];
        pose_log << "\n";
      }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    std::cout << "frames: " << frames << "\t" << "time:" << secs << "s" << std::endl;
    pointCloudGenerator.RequestStop();
    pointCloudGenerator.SavePly(out);
    const vh_stats s = pointCloudGenerator.engine()->stats();
    std::cout << "allocated blocks " << s.allocated_blocks << " arena triangles " << s.arena_triangles << std::endl;
  } catch (const char* msg) {
    std::cerr << "error: " << msg << " (" << vh_last_error() << ")" << std::endl;
    return 1;
  }
  return 0;
}
