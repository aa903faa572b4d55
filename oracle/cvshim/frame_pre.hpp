// frame_pre.hpp -- TEST INFRASTRUCTURE ONLY; force-included (-include) ahead of the reference's src/orbframe.cpp.
// orbframe.cpp relies on <thread> arriving through its include chain and calls one helper of Orbconverter, whose own
// header pulls in Eigen and g2o: the include guard of orbconverter.hpp is pre-defined by the Makefile and the one
// function is declared here (defined in frame_glue.cpp: one header per descriptor row, as orbconverter.cpp:28-36).
#pragma once
#include <array>
#include <climits>
#include <cstdint>
#include <set>
#include <thread>
#include <vector>
#include <opencv2/core/core.hpp>
class Orbconverter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat &Descriptors);
};
