// Drives the drop-in C++ adapter exactly the way OrbFrame does (orbframe.cpp:213-223): construct,
// ExtractFeatures(image, keys, descriptors), read getters and m_vImagePyramid.  Compiled against
// oracle/cvshim (this image has no OpenCV C++ headers).  Reads a raw 8-bit image, writes the
// results as flat binary for the pytest that compares them with the oracle.
#include "orbextractor_b200.hpp"
#include "orbmatcher_b200.hpp"
#include "orbframe_stereo_b200.hpp"
#include "orbvocabulary_b200.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <cstdlib>

// the shim declares these for the reference translation unit; the adapter never calls them
namespace cv {
void resize(const Mat &, Mat &, Size, double, double, int) { abort(); }
void copyMakeBorder(const Mat &, Mat &, int, int, int, int, int) { abort(); }
void FAST(const Mat &, std::vector<KeyPoint> &, int, bool) { abort(); }
void GaussianBlur(const Mat &, Mat &, Size, double, double, int) { abort(); }
float fastAtan2(float, float) { abort(); }
}

// stand-ins for OrbBowVector / OrbFeatureVector (reference include/orbbowvector.hpp:46-50, orbfeaturevector.hpp:37)
struct MockBow {
    std::map<uint32_t, double> m;
    void clear() { m.clear(); }
    void addWeight(uint32_t id, double v) { m[id] += v; }
    void normalize() { double s = 0; for (auto &kv : m) s += std::fabs(kv.second); if (s > 0) for (auto &kv : m) kv.second /= s; }
};
struct MockFeat {
    std::map<uint32_t, std::vector<uint32_t>> m;
    void clear() { m.clear(); }
    void addFeature(uint32_t id, uint32_t i) { m[id].push_back(i); }
};

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s in.raw w h nfeatures nlevels out.bin\n", argv[0]); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]), nl = atoi(argv[5]);
    std::vector<uchar> buf((size_t)w * h);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) { perror("read"); return 2; }
    fclose(f);
    try {
        OrbExtractor ex(nf, 1.2f, nl, 20, 7);
        cv::Mat image(h, w, CV_8UC1, buf.data(), (size_t)w);
        std::vector<cv::KeyPoint> keys;
        cv::Mat desc;
        ex.ExtractFeatures(image, keys, desc);
        ex.ExtractFeatures(image, keys, desc);   // second call on the same instance, like the next frame
        FILE *o = fopen(argv[6], "wb");
        int n = (int)keys.size(), levels = ex.getLevels();
        fwrite(&n, 4, 1, o);
        fwrite(keys.data(), sizeof(cv::KeyPoint), n, o);
        for (int i = 0; i < n; i++) fwrite(desc.ptr(i), 1, 32, o);
        fwrite(&levels, 4, 1, o);
        std::vector<float> sf = ex.getScaleFactors(), is2 = ex.getInverseScaleSigmaSquares();
        fwrite(sf.data(), 4, levels, o);
        fwrite(is2.data(), 4, levels, o);
        // the pyramid member is a view with the reference vector's surface: size, iteration, conversion to std::vector<cv::Mat>&
        {
            long rowsIter = 0, rowsVec = 0;
            for (cv::Mat &m : ex.m_vImagePyramid) rowsIter += m.rows;
            std::vector<cv::Mat> &asVector = ex.m_vImagePyramid;
            for (size_t l = 0; l < asVector.size(); l++) rowsVec += asVector[l].rows;
            if ((int)ex.m_vImagePyramid.size() != levels || rowsIter != rowsVec || rowsIter <= h || ex.m_vImagePyramid.at(0).rows != h ||
                ex.m_vImagePyramid.front().cols != w || ex.m_vImagePyramid.back().rows >= h) {
                fprintf(stderr, "pyramid view: %ld / %ld rows over %zu levels\n", rowsIter, rowsVec, ex.m_vImagePyramid.size());
                return 3;
            }
        }
        for (int l = 0; l < levels; l++) {
            const cv::Mat &m = ex.m_vImagePyramid[l];
            fwrite(&m.cols, 4, 1, o); fwrite(&m.rows, 4, 1, o);
            for (int y = 0; y < m.rows; y++) fwrite(m.ptr(y), 1, m.cols, o);
        }
        // matcher adapter: the descriptors against themselves -> every best match is itself at distance 0
        if (n > 1) {
            orbslam_b200::HammingMatcher hm(n, n);
            hm.SetTrain(desc);
            std::vector<int> idx, d1, d2;
            hm.KnnMatch2(desc, idx, d1, d2);
            fwrite(idx.data(), 4, n, o); fwrite(d1.data(), 4, n, o); fwrite(d2.data(), 4, n, o);
        }
        // stereo adapter: the image against itself -> every match has zero disparity (clamped to 0.01, orbframe.cpp:679-683)
        {
            OrbExtractor exR(nf, 1.2f, nl, 20, 7);
            std::vector<cv::KeyPoint> keysR; cv::Mat descR;
            exR.ExtractFeatures(image, keysR, descR);
            std::vector<float> uR, depth;
            int nm = orbslam_b200::ComputeStereoMatches(ex, exR, 400.0f, 0.0f, uR, depth);
            int nu = (int)uR.size();
            fwrite(&nm, 4, 1, o); fwrite(&nu, 4, 1, o);
            fwrite(uR.data(), 4, nu, o); fwrite(depth.data(), 4, nu, o);
        }
        // vocabulary adapter: a small deterministic 3-ary tree of depth 2 over the frame's own descriptors as node
        // descriptors; bag-of-words bookkeeping as in OrbVocabulary::transform4
        if (n >= 13) {
            const int k = 3, L = 2, nn = 1 + k + k * k;
            std::vector<int32_t> off(nn + 1, 0), ids, wid(nn, -1);
            std::vector<double> wt(nn, 0.0);
            std::vector<uint8_t> nd((size_t)nn * 32);
            for (int v = 0; v < nn; v++) {
                if (v < 1 + k) for (int c = 0; c < k; c++) ids.push_back(1 + v * k + c);
                off[v + 1] = (int32_t)ids.size();
                memcpy(&nd[(size_t)v * 32], desc.ptr(v), 32);
                if (v >= 1 + k) { wid[v] = v - (1 + k); wt[v] = (v % 4 == 0) ? 0.0 : 0.5 + v; }   // some stopped words
            }
            orbslam_b200::VocabularyTransform vt(off, ids, nd, wid, wt, L);
            std::vector<cv::Mat> feats;
            for (int i = 0; i < n; i++) feats.push_back(desc.rowRange(i, i + 1));
            MockBow bow; MockFeat fv;
            vt.transform4(feats, bow, fv, 1);
            int nb = (int)bow.m.size(), nfv = 0;
            for (auto &kv : fv.m) nfv += (int)kv.second.size();
            double sum = 0; for (auto &kv : bow.m) sum += kv.second;
            fwrite(&nb, 4, 1, o); fwrite(&nfv, 4, 1, o); fwrite(&sum, 8, 1, o);
            fwrite(vt.words().data(), 4, n, o); fwrite(vt.nodes().data(), 4, n, o);
            // distinctive descriptors of three "map points" observing rows of this frame
            orbslam_b200::HammingMatcher hm2(4, 4);
            std::vector<int> offs = {0, 5, 5, 12}, idxs = {0, 1, 2, 3, 4, 7, 7, 8, 9, 10, 11, 12}, best;
            hm2.DistinctiveDescriptors(desc, offs, idxs, best);
            fwrite(best.data(), 4, 3, o);
            // SearchByProjection adapter: every second key point of the frame projected 1.25 px to the right of itself
            std::vector<float> uR(n, -1.f), px, py, rad;
            std::vector<unsigned char> occ(n, 0), obs;
            std::vector<int> lvl, pm, asg;
            cv::Mat pd(n / 2, 32, CV_8U);
            for (int i = 0; i + 1 < n; i += 2) {
                memcpy(pd.ptr(i / 2), desc.ptr(i), 32);
                px.push_back(keys[i].pt.x + 1.25f); py.push_back(keys[i].pt.y); lvl.push_back(keys[i].octave);
                rad.push_back(4.0f * sf[keys[i].octave]);
                if (i % 10 == 0) occ[i] = 1;
                obs.push_back((i / 2) % 3 != 0);       // two map points out of three carry observations (orbmatcher.cpp:87-89 after :121)
            }
            int nm = hm2.SearchByProjection(keys, uR, occ, desc, 0.f, 0.f, (float)w, (float)h, pd, px, py, lvl, rad, obs, 0.8f, 100, pm, asg);
            fwrite(&nm, 4, 1, o); fwrite(asg.data(), 4, n, o);
            // GetFeaturesInArea adapter: the same windows at levels [l-1, l]; lists in the reference's order + distances
            std::vector<int> l0(lvl), aoff, aind, adist;
            for (size_t i = 0; i < l0.size(); i++) l0[i] = lvl[i] - 1;
            hm2.AreaDistances(keys, desc, 0.f, 0.f, (float)w, (float)h, pd, px, py, rad, l0, lvl, aoff, aind, adist);
            int total = (int)aind.size(), nq = (int)px.size();
            fwrite(&nq, 4, 1, o); fwrite(&total, 4, 1, o);
            fwrite(aoff.data(), 4, nq + 1, o); fwrite(aind.data(), 4, total, o); fwrite(adist.data(), 4, total, o);
            // AssignFeaturesToGrid adapter into the reference's m_grid shape
            static std::vector<size_t> grid[64][48];
            hm2.AssignFeaturesToGrid(keys, 0.f, 0.f, (float)w, (float)h, grid);
            for (int ix = 0; ix < 64; ix++) for (int iy = 0; iy < 48; iy++) {
                int c = (int)grid[ix][iy].size(); fwrite(&c, 4, 1, o);
                for (size_t v : grid[ix][iy]) { int vi = (int)v; fwrite(&vi, 4, 1, o); }
            }
        }
        // FilterKeyPoints adapter: both extractors' device-resident results, then the stereo matching on what is left
        {
            OrbExtractor exR(nf, 1.2f, nl, 20, 7);
            std::vector<cv::KeyPoint> keysR; cv::Mat descR;
            exR.ExtractFeatures(image, keysR, descR);
            orbslam_b200::FilterKeyPoints(ex, exR, std::array<float, 4>{0.25f * w, 0.75f * w, 0.25f * h, 0.75f * h});
            std::vector<float> uR, depth;
            orbslam_b200::ComputeStereoMatches(ex, exR, 400.0f, 0.0f, uR, depth);
            int nu = (int)uR.size();
            fwrite(&nu, 4, 1, o);
        }
        fclose(o);
    } catch (const std::exception &e) {
        fprintf(stderr, "adapter failed: %s\n", e.what());
        return 1;
    }
    return 0;
}
