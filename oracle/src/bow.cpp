// CPU oracle (test infrastructure only) for the bag-of-words row of SURVEY.md 8(f):
//   BowIndex::transform      bow_index.cpp:59-93     -> DBoW2 TemplatedVocabulary::transform(features, v, fv, levelsUp)
//   BowIndex::add / remove   bow_index.cpp:44-57     inverted lists  std::vector<std::list<MapKf>>
//   BowIndex::getBowSimilar  bow_index.cpp:95-176    words in common through the inverted lists, L1 score, window
// PARITY UNPINNED: DBoW2 is an un-vendored dependency of the reference (absent from /root/reference, version not
// pinned) and the reference holds no test for this path.  The DBoW2 pieces (tree descent, BowVector::addWeight /
// normalize(L1), L1Scoring::score with its lower_bound skipping) are restated from the published library; the index
// and the selection rules follow the cited reference lines statement by statement, with the same containers
// (std::map / std::list / std::sort), so that the CUDA path -- which has no inverted lists at all -- is checked
// against an independent formulation.
#include <algorithm>
#include <cmath>
#include <list>
#include <map>
#include <vector>
#include "common.h"

extern "C" unsigned orc_hamming(const uint32_t *a, const uint32_t *b);   // match.cpp (openvslam/match_base.h:18-39)


// ---- BoW transform ---------------------------------------------------------------------------
// bow_index.cpp:59-93 calls DBoW2's TemplatedVocabulary<ORB>::transform(features, bowVector, featureVector, levelsUp = 4).
// DBoW2 is an un-vendored dependency of the reference (absent from /root/reference, version unpinned): this restates
// its published per-feature descent -- PARITY UNPINNED.  From the root, at every level take the child with the
// smallest Hamming distance (strict '<': the first child wins ties) until a leaf; the word is the leaf's word id and
// weight; the feature-vector node is the node reached at level L - levelsUp (the root when that is <= 0).
// Tree: children of node i are child_ids[child_off[i] .. child_off[i+1]); a leaf has none.
extern "C" void orc_bow_transform(const int *child_off, const int *child_ids, const uint32_t *node_desc, const double *node_weight,
                                  const int *node_word, int n_nodes, int levels, const uint32_t *desc, int n, int levels_up,
                                  int *out_word, double *out_weight, int *out_node) {
    (void)n_nodes;
    const int nid_level = levels - levels_up;
    for (int f = 0; f < n; ++f) {
        int cur = 0, level = 0, nid = 0;
        while (child_off[cur + 1] > child_off[cur]) {
            ++level;
            const int b = child_off[cur], e = child_off[cur + 1];
            int best = child_ids[b];
            unsigned best_d = orc_hamming(desc + 8 * f, node_desc + 8 * (size_t)best);
            for (int c = b + 1; c < e; ++c) {
                const unsigned d = orc_hamming(desc + 8 * f, node_desc + 8 * (size_t)child_ids[c]);
                if (d < best_d) { best_d = d; best = child_ids[c]; }
            }
            cur = best;
            if (level == nid_level) nid = cur;
        }
        out_word[f] = node_word[cur];
        out_weight[f] = node_weight[cur];
        out_node[f] = nid_level <= 0 ? 0 : nid;
    }
}

// ---- BowVector (DBoW2 BowVector.cpp: addWeight, normalize; TemplatedVocabulary::transform, TF-IDF branch) ----------
typedef std::map<unsigned, double> BowVector;

static void add_weight(BowVector &v, unsigned id, double w) {
    BowVector::iterator vit = v.lower_bound(id);
    if (vit != v.end() && !(v.key_comp()(id, vit->first))) vit->second += w;
    else v.insert(vit, BowVector::value_type(id, w));
}

static void normalize_l1(BowVector &v) {
    double norm = 0.0;
    for (BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
    if (norm > 0.0)
        for (BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
}

extern "C" int orc_bow_vector(const int *word, const double *weight, int n, unsigned *out_word, double *out_value) {
    BowVector v;
    for (int f = 0; f < n; ++f)
        if (weight[f] > 0) add_weight(v, (unsigned)word[f], weight[f]);   // "not stopped"
    normalize_l1(v);
    int k = 0;
    for (const auto &e : v) { out_word[k] = e.first; out_value[k] = e.second; ++k; }
    return k;
}

// DBoW2 L1Scoring::score
static double score_l1(const BowVector &v1, const BowVector &v2) {
    BowVector::const_iterator v1_it = v1.begin(), v2_it = v2.begin();
    const BowVector::const_iterator v1_end = v1.end(), v2_end = v2.end();
    double score = 0;
    while (v1_it != v1_end && v2_it != v2_end) {
        const double &vi = v1_it->second, &wi = v2_it->second;
        if (v1_it->first == v2_it->first) {
            score += std::fabs(vi - wi) - std::fabs(vi) - std::fabs(wi);
            ++v1_it;
            ++v2_it;
        } else if (v1_it->first < v2_it->first) {
            v1_it = v1.lower_bound(v2_it->first);
        } else {
            v2_it = v2.lower_bound(v1_it->first);
        }
    }
    return -score / 2.0;
}

// ---- BowIndex (bow_index.cpp:31-57, 95-176) ---------------------------------------------------------------------
namespace {
struct MapKf {
    int mapId, kfId;
    bool operator==(const MapKf &o) const { return mapId == o.mapId && kfId == o.kfId; }                    // :178-180
    bool operator<(const MapKf &o) const { return mapId == o.mapId ? kfId < o.kfId : mapId < o.mapId; }     // :183-188
};
struct BowIndex {
    std::vector<std::list<MapKf>> index;     // one list per word id
    std::map<MapKf, BowVector> bowVec;       // stands in for mapDB.keyframes.at(kfId)->shared->bowVec
};
}  // namespace

extern "C" void *orc_bowindex_create(int vocabulary_size) {
    BowIndex *b = new BowIndex();
    b->index.resize(vocabulary_size);
    return b;
}
extern "C" void orc_bowindex_destroy(void *h) { delete static_cast<BowIndex *>(h); }

extern "C" void orc_bowindex_add(void *h, int map_id, int kf_id, const unsigned *word, const double *value, int n) {
    BowIndex &b = *static_cast<BowIndex *>(h);
    BowVector v;
    for (int i = 0; i < n; ++i) v[word[i]] = value[i];
    for (const auto &w : v) b.index[w.first].push_back(MapKf{map_id, kf_id});
    b.bowVec[MapKf{map_id, kf_id}] = v;
}

extern "C" void orc_bowindex_remove(void *h, int map_id, int kf_id) {
    BowIndex &b = *static_cast<BowIndex *>(h);
    const MapKf mapKf{map_id, kf_id};
    for (auto &l : b.index)
        for (auto it = l.begin(); it != l.end();) {
            if (*it == mapKf) it = l.erase(it);
            else it++;
        }
    b.bowVec.erase(mapKf);
}

extern "C" int orc_bowindex_similar(void *h, const unsigned *q_word, const double *q_value, int nq, int self_map, int self_kf,
                                    float bowMinInCommonRatio, float bowScoreRatio, int *out_map, int *out_kf, float *out_score,
                                    int capacity) {
    BowIndex &b = *static_cast<BowIndex *>(h);
    const MapKf currentMapKf{self_map, self_kf};
    BowVector bowVec;
    for (int i = 0; i < nq; ++i) bowVec[q_word[i]] = q_value[i];

    std::map<MapKf, unsigned int> wordsInCommon;
    for (const auto &pair : bowVec) {
        const unsigned wordId = pair.first;
        if (b.index.at(wordId).empty()) continue;
        const std::list<MapKf> &inNode = b.index.at(wordId);
        for (MapKf mapKf : inNode) {
            if (mapKf == currentMapKf) continue;
            if (!wordsInCommon.count(mapKf)) wordsInCommon[mapKf] = 0;
            ++wordsInCommon.at(mapKf);
        }
    }
    if (wordsInCommon.empty()) return 0;

    unsigned int maxInCommon = 0;
    for (const auto &word : wordsInCommon)
        if (word.second > maxInCommon) maxInCommon = word.second;
    const auto minInCommon = static_cast<unsigned int>(bowMinInCommonRatio * static_cast<float>(maxInCommon));

    struct BowSimilar { MapKf mapKf; float score; };
    std::vector<BowSimilar> similar;
    for (const auto &word : wordsInCommon)
        if (word.second > minInCommon) {
            float score = (float)score_l1(bowVec, b.bowVec.at(word.first));
            similar.push_back(BowSimilar{word.first, score});
        }
    if (similar.empty()) return 0;

    std::sort(similar.begin(), similar.end(), [](const BowSimilar &p1, const BowSimilar &p2) { return p1.score > p2.score; });
    float minScore = similar[0].score * bowScoreRatio;
    auto cut = std::find_if(similar.begin(), similar.end(), [&minScore](const BowSimilar &p) { return p.score < minScore; });
    if (cut != similar.end()) similar.erase(cut, similar.end());
    const int n = (int)similar.size();
    for (int i = 0; i < std::min(n, capacity); ++i) {
        out_map[i] = similar[i].mapKf.mapId;
        out_kf[i] = similar[i].mapKf.kfId;
        out_score[i] = similar[i].score;
    }
    return n;
}
