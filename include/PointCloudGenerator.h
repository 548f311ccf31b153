// PointCloudGenerator.h — drop-in for the reference's facade (/root/reference/include/PointCloudGenerator.h,
// src/PointCloudGenerator.cpp): reads the settings file, owns the engine, turns an RGBDFrame into the raw pointers +
// row-major camera-to-world matrix processFrame takes (PointCloudGenerator.cpp:93-129,139-143), SavePly, Render.
// Header-only. The settings file is the reference's OpenCV-YAML (scene0220_02/scene0220_02.yaml); only the flat
// "Key: value" lines it consumes (PointCloudGenerator.cpp:22-41) are parsed, so OpenCV's FileStorage is not needed.
#ifndef VH_POINTCLOUDGENERATOR_H_
#define VH_POINTCLOUDGENERATOR_H_

#include <chrono>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>

#include "Utils.h"
#include "tsdf.cuh"

namespace ark {

class PointCloudGenerator {
 public:
  explicit PointCloudGenerator(std::string strSettingsFile) : mptRun(nullptr), mpGpuTsdfGenerator(nullptr), mbRequestStop(false) {
    std::map<std::string, double> s = readSettings(strSettingsFile);
    fx_ = (float)s["Camera.fx"]; fy_ = (float)s["Camera.fy"]; cx_ = (float)s["Camera.cx"]; cy_ = (float)s["Camera.cy"];
    width_ = (int)s["Camera.width"]; height_ = (int)s["Camera.height"];
    depthfactor_ = (float)s["DepthMapFactor"]; maxdepth_ = (float)s["MaxDepth"];
    mpGpuTsdfGenerator = new GpuTsdfGenerator(width_, height_, fx_, fy_, cx_, cy_, maxdepth_, (float)s["Voxel.Origin.x"], (float)s["Voxel.Origin.y"],
                                              (float)s["Voxel.Origin.z"], (float)s["Voxel.Size"], (float)s["Voxel.TruncMargin"],
                                              (int)s["Voxel.Dim.x"], (int)s["Voxel.Dim.y"], (int)s["Voxel.Dim.z"]);
    mKeyFrame.frameId = -1;
  }
  ~PointCloudGenerator() {
    RequestStop();
    if (mptRun) { mptRun->join(); delete mptRun; }
    delete mpGpuTsdfGenerator;
  }

  void Start() { mptRun = new std::thread(&PointCloudGenerator::Run, this); }
  void RequestStop() {
    std::unique_lock<std::mutex> lock(mRequestStopMutex);
    mbRequestStop = true;
  }
  bool IsRunning() { std::unique_lock<std::mutex> lock(mRequestStopMutex); return mbRequestStop; }

  // key-frame worker (PointCloudGenerator.cpp:67-91); unlike the reference's loop it sleeps while idle
  void Run() {
    int done = -1;
    for (;;) {
      { std::unique_lock<std::mutex> lock(mRequestStopMutex); if (mbRequestStop) break; }
      RGBDFrame current;
      {
        std::unique_lock<std::mutex> lock(mKeyFrameMutex);
        if (mKeyFrame.frameId != done) { current = mKeyFrame; }
      }
      if (current.frameId == -1 || current.frameId == done) { std::this_thread::sleep_for(std::chrono::milliseconds(1)); continue; }
      done = current.frameId;
      Reproject(current.imRGB, current.imDepth, current.mTcw.inv());
    }
  }

  void OnKeyFrameAvailable(const RGBDFrame& keyFrame) {
    if (mMapRGBDFrame.find(keyFrame.frameId) != mMapRGBDFrame.end()) return;
    std::unique_lock<std::mutex> lock(mKeyFrameMutex);
    mKeyFrame = keyFrame;
    mMapRGBDFrame[keyFrame.frameId] = RGBDFrame();
  }
  void OnFrameAvailable(const RGBDFrame& frame) { std::cout << "OnFrameAvailable" << frame.frameId << std::endl; }
  void OnLoopClosureDetected() { std::cout << "LoopClosureDetected" << std::endl; }

  void SavePly(std::string filename) { mpGpuTsdfGenerator->SavePLY(filename); }
  void Render() { mpGpuTsdfGenerator->render(); }

  // rows 0-2 of Twc, last row 0 0 0 1 (PointCloudGenerator.cpp:118-125)
  void Reproject(const Mat& imRGB, const Mat& imD, const Mat& Twc) {
    float cam2base[16];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) cam2base[r * 4 + c] = Twc.at<float>(r, c);
    cam2base[12] = 0.0f; cam2base[13] = 0.0f; cam2base[14] = 0.0f; cam2base[15] = 1.0f;
    std::memcpy(last_c2w_, cam2base, sizeof(cam2base));
    mpGpuTsdfGenerator->processFrame((float*)imD.datastart, (unsigned char*)imRGB.datastart, cam2base);
  }
  void PushFrame(const RGBDFrame& frame) { Reproject(frame.imRGB, frame.imDepth, frame.mTcw.inv()); }

  GpuTsdfGenerator* engine() { return mpGpuTsdfGenerator; }
  const float* lastPose() const { return last_c2w_; }

 private:
  // "%YAML:1.0" flat mapping: "Key: number" per line, '#' comments
  static std::map<std::string, double> readSettings(const std::string& path) {
    std::map<std::string, double> out;
    std::ifstream f(path);
    if (!f) { std::cerr << "cannot open settings file " << path << std::endl; return out; }
    std::string line;
    while (std::getline(f, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      const size_t colon = line.find(':');
      if (colon == std::string::npos || line[0] == '%') continue;
      std::string key = line.substr(0, colon), val = line.substr(colon + 1);
      key.erase(0, key.find_first_not_of(" \t")); key.erase(key.find_last_not_of(" \t") + 1);
      std::istringstream vs(val);
      double v;
      if (vs >> v) out[key] = v;
    }
    return out;
  }

  std::thread* mptRun;
  GpuTsdfGenerator* mpGpuTsdfGenerator;
  std::map<int, RGBDFrame> mMapRGBDFrame;
  std::mutex mKeyFrameMutex;
  RGBDFrame mKeyFrame;
  std::mutex mFrameMutex;
  RGBDFrame mFrame;
  std::mutex mRequestStopMutex;
  bool mbRequestStop;
  float fx_, fy_, cx_, cy_;
  float maxdepth_;
  int width_, height_;
  float depthfactor_;
  float last_c2w_[16] = {0};
};

}  // namespace ark
#endif  // VH_POINTCLOUDGENERATOR_H_
