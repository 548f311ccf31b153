// Utils.h — drop-in for the reference's include/Utils.h: ark::RGBDFrame {mTcw, imRGB, imDepth, frameId}.
// With OpenCV available (define VH_WITH_OPENCV) the members are cv::Mat as in the reference (Utils.h:15-33); this image
// has no OpenCV C++ headers, so by default ark::Mat is a small owning matrix with the cv::Mat members the hot path's
// callers touch: rows, cols, datastart, at<T>(r, c), copyTo, inv() of a 4x4 (PointCloudGenerator.cpp:118-127,140;
// SaveFrame.cpp:174-206).
#ifndef VH_UTILS_H_
#define VH_UTILS_H_

#ifdef VH_WITH_OPENCV
#include <opencv2/opencv.hpp>
namespace ark { typedef cv::Mat Mat; }
#else
#include <cmath>
#include <cstring>
#include <vector>

namespace ark {

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* datastart = nullptr;

  Mat() {}
  Mat(int r, int c, int channels, int bytes_per_channel) { create(r, c, channels, bytes_per_channel); }
  Mat(const Mat& o) { *this = o; }
  Mat& operator=(const Mat& o) {
    rows = o.rows; cols = o.cols; ch_ = o.ch_; bpc_ = o.bpc_; buf_ = o.buf_;
    datastart = buf_.empty() ? nullptr : buf_.data();
    return *this;
  }
  void create(int r, int c, int channels, int bytes_per_channel) {
    rows = r; cols = c; ch_ = channels; bpc_ = bytes_per_channel;
    buf_.assign((size_t)r * c * channels * bytes_per_channel, 0);
    datastart = buf_.empty() ? nullptr : buf_.data();
  }
  static Mat eye(int n) {
    Mat m(n, n, 1, 4);
    for (int i = 0; i < n; i++) m.at<float>(i, i) = 1.0f;
    return m;
  }
  bool empty() const { return buf_.empty(); }
  int channels() const { return ch_; }
  size_t total_bytes() const { return buf_.size(); }
  template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(datastart)[(size_t)r * cols + c]; }
  template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(datastart)[(size_t)r * cols + c]; }
  void copyTo(Mat& dst) const { dst = *this; }

  // inverse of a square float matrix: Gauss-Jordan with partial pivoting, accumulated in double
  Mat inv() const {
    const int n = rows;
    std::vector<double> a((size_t)n * 2 * n, 0.0);
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) a[(size_t)i * 2 * n + j] = at<float>(i, j); a[(size_t)i * 2 * n + n + i] = 1.0; }
    for (int c = 0; c < n; c++) {
      int p = c;
      for (int r = c + 1; r < n; r++) if (std::fabs(a[(size_t)r * 2 * n + c]) > std::fabs(a[(size_t)p * 2 * n + c])) p = r;
      if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(a[(size_t)p * 2 * n + j], a[(size_t)c * 2 * n + j]);
      const double d = a[(size_t)c * 2 * n + c];
      if (d == 0.0) return Mat(n, n, 1, 4);
      for (int j = 0; j < 2 * n; j++) a[(size_t)c * 2 * n + j] /= d;
      for (int r = 0; r < n; r++) {
        if (r == c) continue;
        const double f = a[(size_t)r * 2 * n + c];
        if (f != 0.0) for (int j = 0; j < 2 * n; j++) a[(size_t)r * 2 * n + j] -= f * a[(size_t)c * 2 * n + j];
      }
    }
    Mat out(n, n, 1, 4);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) out.at<float>(i, j) = (float)a[(size_t)i * 2 * n + n + j];
    return out;
  }

 private:
  int ch_ = 1, bpc_ = 1;
  std::vector<unsigned char> buf_;
};

}  // namespace ark
#endif  // VH_WITH_OPENCV

namespace ark {

class RGBDFrame {
 public:
  Mat mTcw;
  Mat imRGB;
  Mat imDepth;
  int frameId;
  RGBDFrame() : frameId(-1) {
#ifdef VH_WITH_OPENCV
    mTcw = cv::Mat::eye(4, 4, CV_32FC1);
#else
    mTcw = Mat::eye(4);
#endif
  }
  RGBDFrame(const RGBDFrame& frame) {
    frame.mTcw.copyTo(mTcw);
    frame.imRGB.copyTo(imRGB);
    frame.imDepth.copyTo(imDepth);
    frameId = frame.frameId;
  }
  RGBDFrame& operator=(const RGBDFrame& frame) {
    frame.mTcw.copyTo(mTcw);
    frame.imRGB.copyTo(imRGB);
    frame.imDepth.copyTo(imDepth);
    frameId = frame.frameId;
    return *this;
  }
};

}  // namespace ark
#endif  // VH_UTILS_H_
