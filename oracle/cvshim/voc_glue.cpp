// voc_glue.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's own OrbVocabulary (src/orbvocabulary.cpp, orbdescriptor.cpp,
// orbbowvector.cpp, orbfeaturevector.cpp), compiled UNMODIFIED from /root/reference by the `ref` target of
// oracle/Makefile against the cv:: header shim.  include/orbvocabulary.hpp also includes orbmatcher.hpp, which
// drags in the whole SLAM data model; the vocabulary does not use it, so the Makefile pre-defines that header's
// include guard (-DORBMATCHER_H) -- no reference source is touched.
//
// The vocabulary is loaded by the reference's own text loader (loadFromTextFile, :39-118) and run through the
// public transform4 (:168-201), which calls the private transform5 (:203-242) per feature.
#include "orbvocabulary.hpp"

#include <cstdint>
#include <vector>

#ifdef ORBREF_DROPIN_VOC
// libvocdropin.so only: the body of OrbVocabulary::transform4 replaced as INTEGRATION.md section 2c says (the reference's own
// definition in orbvocabulary.o is weakened by the Makefile).  The class cannot get a new member here, so the device-side
// tree of each OrbVocabulary lives in a table keyed by the object; vocref_free drops it.
#include <map>
#include <memory>
#include <mutex>

#include "orbvocabulary_b200.hpp"
static std::mutex g_vocMu;
static std::map<const OrbVocabulary *, std::shared_ptr<orbslam_b200::VocabularyTransform>> g_vocGpu;

void OrbVocabulary::transform4(const std::vector<cv::Mat> features, OrbBowVector &bowVector, OrbFeatureVector &featureVector, int levelsUp) const
{
    std::shared_ptr<orbslam_b200::VocabularyTransform> t;
    {
        std::lock_guard<std::mutex> lk(g_vocMu);
        auto it = g_vocGpu.find(this);
        if (it == g_vocGpu.end()) {
            // the tree as flat arrays: children in m_nodes[v].children order, word_id = -1 for inner nodes
            std::vector<int32_t> off(1, 0), ids, wid;
            std::vector<double> wt;
            std::vector<uint8_t> desc;
            for (const Node &n : m_nodes) {
                ids.insert(ids.end(), n.children.begin(), n.children.end());
                off.push_back((int32_t)ids.size());
                wid.push_back(n.isLeaf() ? (int32_t)n.word_id : -1);
                wt.push_back(n.weight);
                uint8_t row[32] = {0};
                if (!n.descriptor.empty()) memcpy(row, n.descriptor.ptr(0), 32);
                desc.insert(desc.end(), row, row + 32);
            }
            it = g_vocGpu.emplace(this, std::make_shared<orbslam_b200::VocabularyTransform>(off, ids, desc, wid, wt, m_L)).first;
        }
        t = it->second;
    }
    t->transform4(features, bowVector, featureVector, levelsUp);
}
#endif

extern "C" {

void *vocref_load(const char *path) { return new OrbVocabulary(std::string(path)); }
void vocref_free(void *v)
{
#ifdef ORBREF_DROPIN_VOC
    { std::lock_guard<std::mutex> lk(g_vocMu); g_vocGpu.erase((const OrbVocabulary *)v); }
#endif
    delete (OrbVocabulary *)v;
}
int vocref_size(void *v) { return ((OrbVocabulary *)v)->GetSize(); }

static cv::Mat row32(const uint8_t *p)
{
    cv::Mat m(1, 32, CV_8U);
    memcpy(m.ptr(0), p, 32);
    return m;
}

// transform4 on every feature by itself: a one-feature bag of words holds exactly the word (weight normalised to 1)
// and the one-entry feature vector holds the node; both are empty when the word is stopped (weight 0).
void vocref_transform_each(void *v, const uint8_t *feat, int n, int levelsUp, int32_t *word, int32_t *node)
{
    OrbVocabulary *voc = (OrbVocabulary *)v;
    for (int i = 0; i < n; i++) {
        std::vector<cv::Mat> f(1, row32(feat + (size_t)i * 32));
        OrbBowVector bow; OrbFeatureVector fv;
        voc->transform4(f, bow, fv, levelsUp);
        word[i] = bow.empty() ? -1 : (int32_t)bow.begin()->first;
        node[i] = fv.empty() ? -1 : (int32_t)fv.begin()->first;
    }
}

// transform4 on the whole set, as OrbFrame::ComputeBoW calls it: flattened (word, value) and (node, feature) lists
int vocref_transform4(void *v, const uint8_t *feat, int n, int levelsUp, uint32_t *bowIds, double *bowVals, int bowCap, int *nBow,
                      uint32_t *fvNodes, uint32_t *fvFeats, int fvCap, int *nFv)
{
    OrbVocabulary *voc = (OrbVocabulary *)v;
    std::vector<cv::Mat> f;
    for (int i = 0; i < n; i++) f.push_back(row32(feat + (size_t)i * 32));
    OrbBowVector bow; OrbFeatureVector fv;
    voc->transform4(f, bow, fv, levelsUp);
    int a = 0, b = 0;
    for (OrbBowVector::const_iterator it = bow.begin(); it != bow.end(); ++it, ++a)
        if (a < bowCap) { bowIds[a] = it->first; bowVals[a] = it->second; }
    for (OrbFeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t k = 0; k < it->second.size(); k++, b++)
            if (b < fvCap) { fvNodes[b] = it->first; fvFeats[b] = it->second[k]; }
    *nBow = a; *nFv = b;
    return (a <= bowCap && b <= fvCap) ? 0 : -1;
}

} // extern "C"
