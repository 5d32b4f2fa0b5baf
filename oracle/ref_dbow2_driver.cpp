// oracle/ref_dbow2_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's OWN bag-of-words code, compiled unmodified
// where it lies under /root/reference (oracle/Makefile): Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp, BowVector.cpp,
// FeatureVector.cpp, ScoringObject.cpp} + DUtils/Random.cpp, instantiated exactly as include/ORBVocabulary.h does
// (TemplatedVocabulary<FORB::TDescriptor, FORB>).  The vocabulary is read by the reference's own loadFromTextFile from a text file in the
// ORBvoc.txt format that the test writes (ORBvoc.txt itself is not in the checkout); transform() is what
// Frame::ComputeBoW / KeyFrame::ComputeBoW call (levelsup = 4).  Outputs are the std::maps flattened in key order, like
// orc_bow_from_words of oracle/orb_oracle.cpp.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // include/ORBVocabulary.h:27

extern "C" {

void *refv_load(const char *text_file) {
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromTextFile(text_file)) { delete v; return nullptr; }
    return v;
}
void refv_destroy(void *v) { delete (ORBVocabulary *)v; }
int refv_size(void *v) { return (int)((ORBVocabulary *)v)->size(); }

// transform(features, BowVector, FeatureVector, levelsup): returns the BowVector size; word / node per feature through the
// single-feature overload transform(feature, id, weight, nid, levelsup) is protected, so they are recovered from the maps
int refv_transform(void *vp, const uint8_t *feat, int n, int levelsup, int *bow_ids, double *bow_vals, int *fv_nodes, int *fv_ptr, int *fv_idx,
                   int *n_fv_nodes) {
    ORBVocabulary *voc = (ORBVocabulary *)vp;
    std::vector<cv::Mat> features(n);
    for (int i = 0; i < n; i++) {
        features[i].create(1, 32, CV_8U);
        memcpy(features[i].data, feat + (size_t)i * 32, 32);
    }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc->transform(features, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_ids[k] = (int)it->first; bow_vals[k] = it->second; }
    int p = 0, q = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++q) {
        fv_nodes[q] = (int)it->first; fv_ptr[q] = p;
        for (size_t j = 0; j < it->second.size(); j++) fv_idx[p++] = (int)it->second[j];
    }
    fv_ptr[q] = p;
    *n_fv_nodes = q;
    return k;
}

// word of one feature: transform(feature) (TemplatedVocabulary.h:1197-1211)
int refv_word(void *vp, const uint8_t *feat) {
    cv::Mat f;
    f.create(1, 32, CV_8U);
    memcpy(f.data, feat, 32);
    return (int)((ORBVocabulary *)vp)->transform(f);
}

// FORB::distance (FORB.cpp:81-101), the bit-twiddling popcount DBoW2 descends with
int refv_distance(const uint8_t *a, const uint8_t *b) {
    cv::Mat ma, mb;
    ma.create(1, 32, CV_8U); mb.create(1, 32, CV_8U);
    memcpy(ma.data, a, 32); memcpy(mb.data, b, 32);
    return (int)DBoW2::FORB::distance(ma, mb);
}

}  // extern "C"
