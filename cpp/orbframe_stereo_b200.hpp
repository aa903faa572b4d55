// orbframe_stereo_b200.hpp -- OrbFrame::ComputeStereoMatches (reference src/orbframe.cpp:511-705) on the
// device-resident results of the two drop-in extractors.
//
// In the reference's stereo constructor (orbframe.cpp:61-88) the call sequence is
//     ExtractORB(0, left) || ExtractORB(1, right);  CommonSetup();  ComputeStereoMatches();
// Replace the body of ComputeStereoMatches with
//     orbslam_b200::ComputeStereoMatches(*m_ORBextractorLeft, *m_ORBextractorRight, mbf, mb, mvuRight, m_depths);
// The keypoints, descriptors and pyramids of both images are still in HBM from the two ExtractFeatures
// calls, so nothing but the two result vectors crosses PCIe.
// When the frame was given a bounding box, CommonSetup -> UndistortKeyPoints -> FilterKeyPoints (:403-445) has removed key
// points from m_keys / m_descriptors before ComputeStereoMatches runs: call orbslam_b200::FilterKeyPoints(left, right, box)
// in the same place, it removes the same key points from the device-resident results.
#ifndef ORBFRAME_STEREO_B200_HPP
#define ORBFRAME_STEREO_B200_HPP

#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "orbextractor_b200.hpp"

namespace orbslam_b200 {

// OrbFrame::FilterKeyPoints on the device-resident results of both extractors (box = m_boundingBox: x0, x1, y0, y1).
inline void FilterKeyPoints(OrbExtractor &left, OrbExtractor &right, const std::array<float, 4> &box)
{
    if (!left.handle() || !right.handle()) throw std::runtime_error("FilterKeyPoints: extract both images first");
    if (orbx_filter_keypoints(left.handle(), 0, 1, box.data()) != ORBX_OK)
        throw std::runtime_error(std::string("liborbx: ") + orbx_last_error(left.handle()));
    if (orbx_filter_keypoints(right.handle(), 0, 1, box.data()) != ORBX_OK)
        throw std::runtime_error(std::string("liborbx: ") + orbx_last_error(right.handle()));
}

inline int ComputeStereoMatches(OrbExtractor &left, OrbExtractor &right, float mbf, float mb,
                                std::vector<float> &mvuRight, std::vector<float> &mvDepth)
{
    if (!left.handle() || !right.handle()) throw std::runtime_error("ComputeStereoMatches: extract both images first");
    const int cap = orbx_max_keypoints(left.handle());
    mvuRight.assign(cap, -1.0f);
    mvDepth.assign(cap, -1.0f);
    int nLeft = 0, nMatches = 0;
    if (orbx_stereo_match(left.handle(), 0, right.handle(), 0, mbf, mb, mvuRight.data(), mvDepth.data(), cap, &nLeft, &nMatches) != ORBX_OK)
        throw std::runtime_error(std::string("liborbx: ") + orbx_last_error(left.handle()));
    mvuRight.resize(nLeft);
    mvDepth.resize(nLeft);
    return nMatches;
}

}  // namespace orbslam_b200
#endif
