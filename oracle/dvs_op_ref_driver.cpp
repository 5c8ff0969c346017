// C-ABI driver around the reference's own DVSOperator (pydvs/cpp/dvs_op.{hpp,cpp}, compiled unchanged by
// oracle/build_ref.py against oracle/opencv_stub).  Test infrastructure only: it pins oracle/dvs_oracle.c's
// orc_dvs_op_f32 and the CUDA kernel k_step_f32 to reference-built code.
// The driver plays the part of PyDVS::update() (pydvs/cpp/dvs_emu.cpp:126-143): it wraps the caller's planes as
// cv::Mat views and runs the operator over cv::Range(0, rows) -- what cv::parallel_for_ does, serially.
#include "dvs_op.hpp"

extern "C" int ref_dvs_op_f32(float *src, float *diff, float *ref, float *thr, float *ev, float relax, float up,
                              float down, int rows, int cols) {
  if (!src || !diff || !ref || !thr || !ev || rows <= 0 || cols <= 0) return 1;
  const size_t step = static_cast<size_t>(cols) * sizeof(float);
  cv::Mat m_src(rows, cols, src, step), m_diff(rows, cols, diff, step), m_ref(rows, cols, ref, step),
      m_thr(rows, cols, thr, step), m_ev(rows, cols, ev, step);
  DVSOperator op(&m_src, &m_diff, &m_ref, &m_thr, &m_ev, relax, up, down);
  op(cv::Range(0, rows));
  return 0;
}
