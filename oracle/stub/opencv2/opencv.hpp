// Minimal stand-in for the OpenCV types netlib.cpp mentions, so that the reference's
// netlib.cpp compiles unmodified on a box without OpenCV. TEST INFRASTRUCTURE ONLY
// (oracle/_ref build); nothing on the hot path touches cv::Mat.
#ifndef AEFFT_ORACLE_STUB_OPENCV_HPP
#define AEFFT_ORACLE_STUB_OPENCV_HPP
#include <cstddef>
#include <vector>
typedef unsigned char uchar;
namespace cv {
struct Vec3b {
  uchar v[3];
  uchar& operator[](int i) { return v[i]; }
};
struct Mat {
  int rows = 0, cols = 0;
  std::vector<uchar> d;
  template <class T>
  T& at(int r, int c) {
    return *reinterpret_cast<T*>(&d[(std::size_t)(r * cols + c) * sizeof(T)]);
  }
};
}  // namespace cv
#endif
