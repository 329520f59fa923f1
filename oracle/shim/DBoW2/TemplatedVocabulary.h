// Stand-in for DBoW2/TemplatedVocabulary.h (absent dependency; TEST INFRASTRUCTURE ONLY): the published vocabulary
// tree of Galvez-Lopez & Tardos (k-ary tree of descriptors, TF-IDF word weights, L1 scoring) restated with the
// members bow_index.cpp uses: loadFromTextFile (the ORB-SLAM "k L scoring weighting" text format), transform with
// the direct index (levelsup), score, size.  loadFromBinaryFile (.dbow2, a fork-specific format) is not restated.
#pragma once
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>
#include "BowVector.h"
#include "FeatureVector.h"
namespace DBoW2 {
template <class TDescriptor, class F> class TemplatedVocabulary {
public:
    struct Node {
        NodeId id = 0, parent = 0;
        WordValue weight = 0;
        std::vector<NodeId> children;
        TDescriptor descriptor;
        WordId word_id = 0;
        bool isLeaf() const { return children.empty(); }
    };
    int m_k = 10, m_L = 5;
    WeightingType m_weighting = TF_IDF;
    ScoringType m_scoring = L1_NORM;
    std::vector<Node> m_nodes;
    std::vector<NodeId> m_words;   // word id -> node id

    unsigned int size() const { return (unsigned)m_words.size(); }
    bool empty() const { return m_words.empty(); }

    // C stdio on purpose: this toolchain links libstdc++ statically into every .so, and iostream objects of a second
    // libstdc++ copy inside a Python process (numpy loads the shared one) are not safe to use
    bool loadFromTextFile(const std::string &filename) {
        FILE *f = std::fopen(filename.c_str(), "r");
        if (!f) return false;
        m_words.clear(); m_nodes.clear();
        int n1 = 0, n2 = 0;
        if (std::fscanf(f, "%d %d %d %d", &m_k, &m_L, &n1, &n2) != 4 || m_k < 0 || m_k > 20 || m_L < 1 || m_L > 10 || n1 < 0
            || n1 > 5 || n2 < 0 || n2 > 3) { std::fclose(f); return false; }
        m_scoring = (ScoringType)n1;
        m_weighting = (WeightingType)n2;
        m_nodes.resize(1);
        m_nodes[0].id = 0;
        int pid, is_leaf;
        while (std::fscanf(f, "%d %d", &pid, &is_leaf) == 2) {
            const NodeId nid = (NodeId)m_nodes.size();
            m_nodes.resize(nid + 1);
            m_nodes[nid].id = nid;
            m_nodes[nid].parent = (NodeId)pid;
            m_nodes[pid].children.push_back(nid);
            m_nodes[nid].descriptor = cv::Mat(1, F::L, CV_8U);
            for (int i = 0; i < F::L; ++i) { int v = 0; if (std::fscanf(f, "%d", &v) != 1) { std::fclose(f); return false; } m_nodes[nid].descriptor.data[i] = (unsigned char)v; }
            double w = 0;
            if (std::fscanf(f, "%lf", &w) != 1) { std::fclose(f); return false; }
            m_nodes[nid].weight = w;
            if (is_leaf > 0) { m_nodes[nid].word_id = (WordId)m_words.size(); m_words.push_back(nid); }
        }
        std::fclose(f);
        return true;
    }
    void loadFromBinaryFile(const std::string &) { std::abort(); }

    // tree descent of one feature
    void transform(const TDescriptor &feature, WordId &word_id, WordValue &weight, NodeId *nid, int levelsup) const {
        const int nid_level = m_L - levelsup;
        if (nid_level <= 0 && nid != nullptr) *nid = 0;
        NodeId final_id = 0;
        int current_level = 0;
        do {
            ++current_level;
            const std::vector<NodeId> &nodes = m_nodes[final_id].children;
            final_id = nodes[0];
            double best_d = F::distance(feature, m_nodes[final_id].descriptor);
            for (auto nit = nodes.begin() + 1; nit != nodes.end(); ++nit) {
                const NodeId id = *nit;
                const double d = F::distance(feature, m_nodes[id].descriptor);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (nid != nullptr && current_level == nid_level) *nid = final_id;
        } while (!m_nodes[final_id].isLeaf());
        word_id = m_nodes[final_id].word_id;
        weight = m_nodes[final_id].weight;
    }
    void transform(const std::vector<TDescriptor> &features, BowVector &v, FeatureVector &fv, int levelsup) const {
        v.clear();
        fv.clear();
        if (empty()) return;
        const bool must = m_scoring == L1_NORM || m_scoring == L2_NORM || m_scoring == CHI_SQUARE || m_scoring == KL
                          || m_scoring == BHATTACHARYYA;   // GeneralScoring::mustNormalize (DOT_PRODUCT: no)
        const LNorm norm = m_scoring == L2_NORM ? L2 : L1;
        unsigned int i_feature = 0;
        for (auto fit = features.begin(); fit < features.end(); ++fit, ++i_feature) {
            WordId id; NodeId nid; WordValue w;
            transform(*fit, id, w, &nid, levelsup);
            if (w > 0) {
                if (m_weighting == TF || m_weighting == TF_IDF) v.addWeight(id, w); else v.addIfNotExist(id, w);
                fv.addFeature(nid, i_feature);
            }
        }
        if (!v.empty() && !must && (m_weighting == TF || m_weighting == TF_IDF)) {
            const double nd = (double)v.size();
            for (auto &e : v) e.second /= nd;
        }
        if (must) v.normalize(norm);
    }
    // L1Scoring::score
    double score(const BowVector &v1, const BowVector &v2) const {
        auto v1_it = v1.begin(), v2_it = v2.begin();
        const auto v1_end = v1.end(), v2_end = v2.end();
        double score = 0;
        while (v1_it != v1_end && v2_it != v2_end) {
            const WordValue &vi = v1_it->second, &wi = v2_it->second;
            if (v1_it->first == v2_it->first) { score += std::fabs(vi - wi) - std::fabs(vi) - std::fabs(wi); ++v1_it; ++v2_it; }
            else if (v1_it->first < v2_it->first) v1_it = v1.lower_bound(v2_it->first);
            else v2_it = v2.lower_bound(v1_it->first);
        }
        return -score / 2.0;
    }
};
}  // namespace DBoW2
