// header shim -- TEST INFRASTRUCTURE ONLY: the reference headers include calib3d.  The one function used
// (cv::undistortPoints, orbframe.cpp:467,494) is only reached with non-zero distortion coefficients, which the
// oracle build never supplies: it aborts instead of pretending.
#pragma once
#include <opencv2/core/core.hpp>
#include <cstdlib>
namespace cv {
inline void undistortPoints(const Mat &, Mat &, const Mat &, const Mat &, const Mat &, const Mat &) { abort(); }
}
