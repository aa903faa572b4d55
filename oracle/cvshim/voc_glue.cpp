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

extern "C" {

void *vocref_load(const char *path) { return new OrbVocabulary(std::string(path)); }
void vocref_free(void *v) { delete (OrbVocabulary *)v; }
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
