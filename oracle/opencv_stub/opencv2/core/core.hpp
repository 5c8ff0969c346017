// Minimal stand-in for the four OpenCV headers pydvs/cpp/dvs_op.hpp includes (test infrastructure only).
// OpenCV's C++ headers are not installed in this image; the reference's DVSOperator touches exactly this much of
// the API: cv::Mat::ptr<T>(row), cv::Mat::cols, cv::Range::start / end and the cv::ParallelLoopBody base class.
// With these declarations /root/reference/pydvs/cpp/dvs_op.cpp compiles UNCHANGED (oracle/build_ref.py).
#ifndef PYDVS_ORACLE_OPENCV_STUB_CORE_HPP
#define PYDVS_ORACLE_OPENCV_STUB_CORE_HPP
#include <cstddef>
#include <cstdint>

namespace cv {

class Range {
 public:
  Range() : start(0), end(0) {}
  Range(int _start, int _end) : start(_start), end(_end) {}
  int start, end;
};

class ParallelLoopBody {
 public:
  virtual ~ParallelLoopBody() {}
  virtual void operator()(const Range &range) const = 0;
};

// A non-owning 2-D view: rows x cols elements, `step` bytes between rows (cv::Mat's memory model).
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), step(0) {}
  Mat(int _rows, int _cols, void *_data, size_t _step)
      : rows(_rows), cols(_cols), data(static_cast<unsigned char *>(_data)), step(_step) {}
  template <typename T> T *ptr(int row = 0) { return reinterpret_cast<T *>(data + static_cast<size_t>(row) * step); }
  template <typename T> const T *ptr(int row = 0) const {
    return reinterpret_cast<const T *>(data + static_cast<size_t>(row) * step);
  }
  int rows, cols;
  unsigned char *data;
  size_t step;
};

}  // namespace cv
#endif
