// Minimal stand-in for the handful of OpenCV types the Rover-SLAM front-end class surface uses
// (cv::Mat, cv::KeyPoint, cv::Point2f, cv::InputArray).  ONLY for building/testing the host classes in
// an image without OpenCV (-DROVER_FE_OPENCV_SHIM); with real OpenCV 3.4/4.x on the include path this file
// is not used and the same sources compile against it.
#pragma once
#include <stdint.h>
#include <string.h>

#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
  Point_ operator/(T s) const { return Point_(x / s, y / s); }
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;

struct Size {
  int width, height;
  Size(int w = 0, int h = 0) : width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(float x, float y, float sz, float ang = -1, float resp = 0, int oct = 0, int cid = -1)
      : pt(x, y), size(sz), angle(ang), response(resp), octave(oct), class_id(cid) {}
};

class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), data(static_cast<uint8_t*>(ext)), type_(type) {
    step = stp ? stp : static_cast<size_t>(c) * elemSize();
  }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = static_cast<size_t>(c) * elemSize();
    buf_.reset(new uint8_t[static_cast<size_t>(r) * step + 16], std::default_delete<uint8_t[]>());
    data = buf_.get();
  }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize() const { return (depth() == CV_32F ? 4 : 1) * static_cast<size_t>(channels()); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == static_cast<size_t>(cols) * elemSize(); }
  Size size() const { return Size(cols, rows); }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + static_cast<size_t>(r) * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + static_cast<size_t>(r) * step); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; ++r) memcpy(m.data + r * m.step, data + r * step, m.step);
    return m;
  }
  void convertTo(Mat& dst, int rtype, double alpha = 1.0) const {
    if (depth() != CV_8U || (rtype & 7) != CV_32F) throw std::invalid_argument("shim convertTo: only 8U->32F");
    Mat out(rows, cols, CV_MAKETYPE(CV_32F, channels()));
    const float a = static_cast<float>(alpha);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols * channels(); ++c) out.ptr<float>(r)[c] = static_cast<float>(ptr<uint8_t>(r)[c]) * a;
    dst = out;
  }
 private:
  int type_ = 0;
  std::shared_ptr<uint8_t> buf_;
};

enum ColorConversionCodes { COLOR_BGR2RGB = 4, COLOR_RGB2GRAY = 7 };
// the two conversions transform.cpp uses; RGB2GRAY with OpenCV's weights (0.299, 0.587, 0.114)
inline void cvtColor(const Mat& src, Mat& dst, int code) {
  if (src.channels() != 3) throw std::invalid_argument("shim cvtColor: 3-channel input expected");
  const bool f32 = src.depth() == CV_32F;
  if (code == COLOR_BGR2RGB) {
    Mat out(src.rows, src.cols, src.type());
    for (int r = 0; r < src.rows; ++r)
      for (int c = 0; c < src.cols; ++c)
        for (int k = 0; k < 3; ++k) {
          if (f32) out.ptr<float>(r)[3 * c + k] = src.ptr<float>(r)[3 * c + 2 - k];
          else out.ptr<uint8_t>(r)[3 * c + k] = src.ptr<uint8_t>(r)[3 * c + 2 - k];
        }
    dst = out;
  } else if (code == COLOR_RGB2GRAY) {
    Mat out(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
    for (int r = 0; r < src.rows; ++r)
      for (int c = 0; c < src.cols; ++c) {
        if (f32) {
          const float* p = src.ptr<float>(r) + 3 * c;
          out.ptr<float>(r)[c] = p[0] * 0.299f + p[1] * 0.587f + p[2] * 0.114f;
        } else {
          const uint8_t* p = src.ptr<uint8_t>(r) + 3 * c;
          out.ptr<uint8_t>(r)[c] = static_cast<uint8_t>((p[0] * 4899 + p[1] * 9617 + p[2] * 1868 + 8192) >> 14);
        }
      }
    dst = out;
  } else {
    throw std::invalid_argument("shim cvtColor: unsupported code");
  }
}

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(m) {}
  Mat getMat() const { return m_; }
  bool empty() const { return m_.empty(); }
 private:
  Mat m_;
};
typedef const _InputArray& InputArray;

}  // namespace cv
