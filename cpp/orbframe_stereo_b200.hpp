// orbframe_stereo_b200.hpp -- OrbFrame::ComputeStereoMatches (reference src/orbframe.cpp:511-705) on the
// device-resident results of the two drop-in extractors.
//
// In the reference's stereo constructor (orbframe.cpp:61-88) the call sequence is
//     ExtractORB(0, left) || ExtractORB(1, right);  CommonSetup();  ComputeStereoMatches();
// Replace the body of ComputeStereoMatches with
//     orbslam_b200::ComputeStereoMatches(*m_ORBextractorLeft, *m_ORBextractorRight, mbf, mb, mvuRight, m_depths);
// The keypoints, descriptors and pyramids of both images are still in HBM from the two ExtractFeatures
// calls, so nothing but the two result vectors crosses PCIe.  (m_keys / m_keysRight must be the
// extractors' unfiltered output, as they are at that point of the constructor.)
#ifndef ORBFRAME_STEREO_B200_HPP
#define ORBFRAME_STEREO_B200_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include "orbextractor_b200.hpp"

namespace orbslam_b200 {

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
