// SaveFrame.h — drop-in for the reference's frame I/O (/root/reference/include/SaveFrame.h, src/SaveFrame.cpp): same
// class, same on-disk layout <folder>/RGB/<id>.*, <folder>/depth/<id>.png (u16 millimetres), <folder>/tcw/<id>.txt
// (4x4 camera-to-world text; frameLoad stores its inverse in mTcw, SaveFrame.cpp:198-206). Header-only; OpenCV is
// optional: without it PNG/PPM are decoded by include/vh_image_io.h (zlib) and JPEG colour images are reported missing.
#ifndef VH_SAVEFRAME_H_
#define VH_SAVEFRAME_H_

#include <sys/stat.h>
#include <sys/types.h>

#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <thread>

#include "Utils.h"
#include "vh_image_io.h"

namespace ark {

class SaveFrame {
 public:
  // the reference exits when a folder is missing (SaveFrame.cpp:14-29); so does this, with the same message
  explicit SaveFrame(std::string folder) : mptRun(nullptr), folderPath(folder), mbRequestStop(false) {
    requireFolder(folderPath);
    rgbPath = folderPath + "RGB/";
    depthPath = folderPath + "depth/";
    tcwPath = folderPath + "tcw/";
    requireFolder(rgbPath);
    requireFolder(depthPath);
    requireFolder(tcwPath);
    mKeyFrame.frameId = -1;
  }

  void Start() { mptRun = new std::thread(&SaveFrame::Run, this); }
  void RequestStop() { std::unique_lock<std::mutex> lock(mRequestStopMutex); mbRequestStop = true; }
  bool IsRunning() { std::unique_lock<std::mutex> lock(mRequestStopMutex); return mbRequestStop; }
  void Run() {}
  void OnKeyFrameAvailable(const RGBDFrame& keyFrame) {
    if (mMapRGBDFrame.find(keyFrame.frameId) != mMapRGBDFrame.end()) return;
    std::cout << "OnKeyFrameAvailable" << keyFrame.frameId << std::endl;
    mKeyFrame = keyFrame;
    mMapRGBDFrame[keyFrame.frameId] = RGBDFrame();
  }
  void OnFrameAvailable(const RGBDFrame& frame) { std::cout << "OnFrameAvailable" << frame.frameId << std::endl; }
  void OnLoopClosureDetected() { std::cout << "LoopClosureDetected" << std::endl; }

  // RGB as 8-bit PNG, depth as 16-bit PNG in millimetres, pose as text (the reference writes an OpenCV XML here,
  // SaveFrame.cpp:139-141; the text form is what frameLoad reads back)
  void frameWrite(const RGBDFrame& frame) {
    if (mMapRGBDFrame.find(frame.frameId) != mMapRGBDFrame.end()) return;
    std::cout << "frameWrite frame = " << frame.frameId << std::endl;
    const std::string id = std::to_string(frame.frameId);
    vhio::Raster rgb; rgb.width = frame.imRGB.cols; rgb.height = frame.imRGB.rows; rgb.channels = 3; rgb.bits = 8;
    rgb.data.assign(frame.imRGB.datastart, frame.imRGB.datastart + (size_t)rgb.width * rgb.height * 3);
    vhio::write_png(rgbPath + id + ".png", rgb);
    vhio::Raster d; d.width = frame.imDepth.cols; d.height = frame.imDepth.rows; d.channels = 1; d.bits = 16;
    d.data.resize((size_t)d.width * d.height * 2);
    const float* src = reinterpret_cast<const float*>(frame.imDepth.datastart);
    for (size_t i = 0; i < (size_t)d.width * d.height; i++) {
      const float mm = src[i] * 1000.0f;                                     // convertTo(CV_16UC1, 1000): round, saturate
      reinterpret_cast<uint16_t*>(d.data.data())[i] = (uint16_t)(mm <= 0.f ? 0 : mm >= 65535.f ? 65535 : (int)std::lrintf(mm));
    }
    vhio::write_png(depthPath + id + ".png", d);
    const Mat c2w = frame.mTcw.inv();
    std::ofstream t(tcwPath + id + ".txt");
    t.precision(9);
    for (int i = 0; i < 4; i++) { for (int k = 0; k < 4; k++) t << c2w.at<float>(i, k) << (k == 3 ? "\n" : " "); }
    mMapRGBDFrame[frame.frameId] = RGBDFrame();
  }

  RGBDFrame frameLoad(int frameId) {
    std::cout << "frameLoad frame ==================== " << frameId << std::endl;
    RGBDFrame frame;
    frame.frameId = frameId;
    const std::string id = std::to_string(frameId);
    vhio::Raster rgb;
    if (!vhio::read_png(rgbPath + id + ".png", rgb) && !vhio::read_pnm(rgbPath + id + ".ppm", rgb)) {
      frame.frameId = -1;                                                   // missing frame, as SaveFrame.cpp:166-169
      return frame;
    }
    if (rgb.bits != 8 || rgb.channels < 3) { frame.frameId = -1; return frame; }
    if (rgb.channels == 4) {                                                // drop alpha
      vhio::Raster t = rgb; t.channels = 3; t.data.resize((size_t)rgb.width * rgb.height * 3);
      for (size_t i = 0; i < (size_t)rgb.width * rgb.height; i++) std::memcpy(&t.data[3 * i], &rgb.data[4 * i], 3);
      rgb = t;
    }
    rgb = vhio::resize_nearest(rgb, 640, 480);                              // cv::resize(rgbBig, imRGB, Size(640,480)), SaveFrame.cpp:171
    frame.imRGB.create(480, 640, 3, 1);
    std::memcpy(frame.imRGB.datastart, rgb.data.data(), rgb.data.size());
    vhio::Raster d;
    if ((!vhio::read_png(depthPath + id + ".png", d) && !vhio::read_pnm(depthPath + id + ".pgm", d)) || d.bits != 16 || d.channels != 1) {
      frame.frameId = -1;
      return frame;
    }
    frame.imDepth.create(d.height, d.width, 1, 4);
    float* dst = reinterpret_cast<float*>(frame.imDepth.datastart);
    const uint16_t* mm = reinterpret_cast<const uint16_t*>(d.data.data());
    for (size_t i = 0; i < (size_t)d.width * d.height; i++) dst[i] = (float)((double)(float)mm[i] * 0.001);   // convertTo(CV_32FC1); *= 0.001
    Mat tcw = Mat::eye(4);
    std::ifstream tcwFile(tcwPath + id + ".txt");
    for (int i = 0; i < 4; ++i)
      for (int k = 0; k < 4; ++k) tcwFile >> tcw.at<float>(i, k);
    frame.mTcw = tcw.inv();
    return frame;
  }

 private:
  static void requireFolder(const std::string& p) {
    struct stat info;
    if (stat(p.c_str(), &info) != 0) { std::cout << "Error:" << p << " doesn't exist!" << std::endl; exit(1); }
    if (info.st_mode & S_IFDIR) std::cout << p << " is a directory" << std::endl;
    else std::cout << p << " is no directory" << std::endl;
  }

  std::thread* mptRun;
  std::string folderPath, rgbPath, depthPath, tcwPath, depth_to_tcw_Path;
  std::map<int, RGBDFrame> mMapRGBDFrame;
  std::mutex mKeyFrameMutex;
  RGBDFrame mKeyFrame;
  std::mutex mFrameMutex;
  RGBDFrame mFrame;
  std::mutex mRequestStopMutex;
  bool mbRequestStop;
};

}  // namespace ark
#endif  // VH_SAVEFRAME_H_
