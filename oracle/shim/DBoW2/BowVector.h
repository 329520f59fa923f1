// Stand-in for DBoW2/BowVector.h (DBoW2 is an un-vendored dependency of the reference, absent offline; TEST
// INFRASTRUCTURE ONLY).  Restates the published container: std::map<WordId, WordValue> with addWeight / normalize.
#pragma once
#include <cmath>
#include <map>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
enum LNorm { L1, L2 };
enum WeightingType { TF_IDF, TF, IDF, BINARY };
enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };
class BowVector : public std::map<WordId, WordValue> {
public:
    void addWeight(WordId id, WordValue v) {
        auto vit = this->lower_bound(id);
        if (vit != this->end() && !(this->key_comp()(id, vit->first))) vit->second += v;
        else this->insert(vit, value_type(id, v));
    }
    void addIfNotExist(WordId id, WordValue v) {
        auto vit = this->lower_bound(id);
        if (vit == this->end() || (this->key_comp()(id, vit->first))) this->insert(vit, value_type(id, v));
    }
    void normalize(LNorm norm_type) {
        double norm = 0.0;
        if (norm_type == L1) { for (auto &e : *this) norm += std::fabs(e.second); }
        else { for (auto &e : *this) norm += e.second * e.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto &e : *this) e.second /= norm;
    }
};
}  // namespace DBoW2
