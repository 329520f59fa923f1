// ORACLE (test infrastructure only): Hamming distance, the brute-force degenerate case of
// matchForLoopClosures, and the angle-consistency histogram.
// Follows openvslam/match_base.h:18-39, keyframe_matcher.cpp:50-158, and
// openvslam/match_angle_checker.h:61-134.
#include "common.h"
#include <cmath>
#include <map>
#include <numeric>

namespace orc {

static inline unsigned hamming(const uint32_t *pa, const uint32_t *pb) {
    constexpr uint32_t m1 = 0x55555555U, m2 = 0x33333333U, m3 = 0x0F0F0F0FU, m4 = 0x01010101U;
    unsigned dist = 0;
    for (unsigned i = 0; i < 8; ++i) {
        uint32_t v = pa[i] ^ pb[i];
        v -= ((v >> 1) & m1);
        v = (v & m2) + ((v >> 2) & m2);
        dist += (((v + (v >> 4)) & m3) * m4) >> 24;
    }
    return dist;
}

struct AngleChecker {  // angle_checker<int>(30, 3)
    static constexpr unsigned LEN = 30, NUM_VALID = 3;
    const float inv_len = 1.0f / LEN;
    std::vector<std::vector<int>> hist{LEN};
    static int bin_of(float delta, float inv_len) {
        if (delta < 0.0) delta += 360.0;   // double constants in the reference: float -> double -> float
        if (360.0 <= delta) delta -= 360.0;
        return cv_round(delta * inv_len);
    }
    void append(float delta, int id) { hist.at((unsigned)bin_of(delta, inv_len)).push_back(id); }
    std::vector<unsigned> order() const {
        std::vector<unsigned> idx(LEN);
        std::iota(idx.begin(), idx.end(), 0);
        std::sort(idx.begin(), idx.end(),  // unstable, exactly as the reference calls it
                  [this](const unsigned a, const unsigned b) { return hist.at(a).size() > hist.at(b).size(); });
        return idx;
    }
    std::vector<int> invalid() const {
        std::vector<int> out;
        const auto bins = order();
        for (unsigned b = 0; b < LEN; ++b) {
            const bool valid = std::any_of(bins.begin(), bins.begin() + NUM_VALID, [b](unsigned i) { return b == i; });
            if (!valid) out.insert(out.end(), hist[b].begin(), hist[b].end());
        }
        return out;
    }
};

unsigned match_bruteforce(const uint32_t *dA, const float *aA, int nA, const uint32_t *dB, const float *aB, int nB,
                          float ratio, unsigned thr, bool check_orientation, bool ratio_is_double, int *matches) {
    constexpr unsigned MAX_DIST = 256;
    unsigned num = 0;
    AngleChecker checker;
    for (int i = 0; i < nA; ++i) matches[i] = -1;
    std::vector<bool> taken(nB, false);
    for (int i1 = 0; i1 < nA; ++i1) {
        unsigned best = MAX_DIST, second = MAX_DIST;
        int best_idx = -1;
        for (int i2 = 0; i2 < nB; ++i2) {
            if (taken[i2]) continue;
            const unsigned d = hamming(dA + 8 * i1, dB + 8 * i2);
            if (d < best) { second = best; best = d; best_idx = i2; }
            else if (d < second) second = d;
        }
        if (thr < best) continue;
        if (ratio_is_double ? ((double)ratio * second < (double)static_cast<float>(best))
                            : (ratio * second < static_cast<float>(best))) continue;
        matches[i1] = best_idx;
        taken[best_idx] = true;
        ++num;
        if (check_orientation) checker.append(aA[i1] - aB[best_idx], i1);
    }
    if (check_orientation)
        for (int idx : checker.invalid()) { matches[idx] = -1; --num; }
    return num;
}

// matchForLoopClosures as the reference runs it (keyframe_matcher.cpp:50-158): the DBoW2 feature vectors
// (std::map<NodeId, std::vector<unsigned>>, indices in insertion order) restrict the comparison to features
// under the same vocabulary node; the two maps are walked in node order with lower_bound skips (:70-146).
// nodeA / nodeB give the node of every feature (-1: in no node); eligA / eligB are the map-point filters of
// :79-84 and :93-96 evaluated by the caller (nullptr: every feature passes).
unsigned match_bow(const uint32_t *dA, const float *aA, const int *nodeA, const unsigned char *eligA, int nA,
                   const uint32_t *dB, const float *aB, const int *nodeB, const unsigned char *eligB, int nB,
                   float ratio, unsigned thr, bool check_orientation, bool ratio_is_double, int *matches) {
    constexpr unsigned MAX_DIST = 256;
    std::map<int, std::vector<unsigned>> fv1, fv2;
    for (int i = 0; i < nA; ++i) if (nodeA[i] >= 0) fv1[nodeA[i]].push_back(i);
    for (int i = 0; i < nB; ++i) if (nodeB[i] >= 0) fv2[nodeB[i]].push_back(i);
    unsigned num = 0;
    AngleChecker checker;
    for (int i = 0; i < nA; ++i) matches[i] = -1;
    std::vector<bool> taken(nB, false);
    auto it1 = fv1.begin();
    auto it2 = fv2.begin();
    while (it1 != fv1.end() && it2 != fv2.end()) {
        if (it1->first == it2->first) {
            for (const auto i1 : it1->second) {
                if (eligA && !eligA[i1]) continue;
                unsigned best = MAX_DIST, second = MAX_DIST;
                int best_idx = -1;
                for (const auto i2 : it2->second) {
                    if (eligB && !eligB[i2]) continue;
                    if (taken[i2]) continue;
                    const unsigned d = hamming(dA + 8 * i1, dB + 8 * i2);
                    if (d < best) { second = best; best = d; best_idx = (int)i2; }
                    else if (d < second) second = d;
                }
                if (thr < best) continue;
                if (ratio_is_double ? ((double)ratio * second < (double)static_cast<float>(best))
                                    : (ratio * second < static_cast<float>(best))) continue;
                matches[i1] = best_idx;
                taken[best_idx] = true;
                ++num;
                if (check_orientation) checker.append(aA[i1] - aB[best_idx], (int)i1);
            }
            ++it1;
            ++it2;
        } else if (it1->first < it2->first) it1 = fv1.lower_bound(it2->first);
        else it2 = fv2.lower_bound(it1->first);
    }
    if (check_orientation)
        for (int idx : checker.invalid()) { matches[idx] = -1; --num; }
    return num;
}

// check_epipolar_constraint (keyframe_matcher.cpp:23-44) on plain doubles.  E is row-major.  Eigen evaluates a fixed
// size-3 sum as p0 + (p1 + p2) (redux_novec_unroller halves the range) [recall: Eigen is not in the reference tree].
static inline double sum3(double p0, double p1, double p2) { return p0 + (p1 + p2); }
bool check_epipolar(const double *b1, const double *b2, const double *E, float scale, float residual_deg_thr) {
    const double e[3] = {sum3(E[0] * b2[0], E[1] * b2[1], E[2] * b2[2]), sum3(E[3] * b2[0], E[4] * b2[1], E[5] * b2[2]),
                         sum3(E[6] * b2[0], E[7] * b2[1], E[8] * b2[2])};
    const double cos_residual = sum3(e[0] * b1[0], e[1] * b1[1], e[2] * b1[2]) / std::sqrt(sum3(e[0] * e[0], e[1] * e[1], e[2] * e[2]));
    const double residual_rad = M_PI / 2.0 - std::abs(std::acos(cos_residual));
    const double residual_rad_thr = residual_deg_thr * M_PI / 180.0;
    return residual_rad < residual_rad_thr * scale;
}

// matchForTriangulationDBoW (keyframe_matcher.cpp:160-293): features WITHOUT a map point (elig = 1), node buckets as in
// match_bow, per kf1 feature the last kf2 feature with distance <= thr and <= the best so far that passes the epipolar
// test, uniqueness in kf2, angle histogram over all nodes.
unsigned match_triangulation(const uint32_t *dA, const float *aA, const int *octA, const double *bearA, const int *nodeA,
                             const unsigned char *eligA, int nA, const uint32_t *dB, const float *aB, const double *bearB,
                             const int *nodeB, const unsigned char *eligB, int nB, const double *E, const float *scale_factors,
                             float residual_deg_thr, unsigned thr, bool check_orientation, int *matches) {
    std::map<int, std::vector<unsigned>> fv1, fv2;
    for (int i = 0; i < nA; ++i) if (nodeA[i] >= 0) fv1[nodeA[i]].push_back(i);
    for (int i = 0; i < nB; ++i) if (nodeB[i] >= 0) fv2[nodeB[i]].push_back(i);
    unsigned num = 0;
    AngleChecker checker;
    for (int i = 0; i < nA; ++i) matches[i] = -1;
    std::vector<bool> taken(nB, false);
    auto it1 = fv1.begin();
    auto it2 = fv2.begin();
    while (it1 != fv1.end() && it2 != fv2.end()) {
        if (it1->first == it2->first) {
            for (const auto i1 : it1->second) {
                if (eligA && !eligA[i1]) continue;
                unsigned best = thr;
                int best_idx = -1;
                for (const auto i2 : it2->second) {
                    if (eligB && !eligB[i2]) continue;
                    if (taken[i2]) continue;
                    const unsigned d = hamming(dA + 8 * i1, dB + 8 * i2);
                    if (d > thr || d > best) continue;
                    if (check_epipolar(bearA + 3 * i1, bearB + 3 * i2, E, scale_factors[octA[i1]], residual_deg_thr)) {
                        best_idx = (int)i2;
                        best = d;
                    }
                }
                if (best_idx < 0) continue;
                taken[best_idx] = true;
                matches[i1] = best_idx;
                ++num;
                if (check_orientation) checker.append(aA[i1] - aB[best_idx], (int)i1);
            }
            ++it1;
            ++it2;
        } else if (it1->first < it2->first) it1 = fv1.lower_bound(it2->first);
        else it2 = fv2.lower_bound(it1->first);
    }
    if (check_orientation)
        for (int idx : checker.invalid()) { matches[idx] = -1; --num; }
    return num;
}

}  // namespace orc

using namespace orc;

extern "C" unsigned orc_hamming(const uint32_t *a, const uint32_t *b) { return hamming(a, b); }
extern "C" unsigned orc_match_bruteforce(const uint32_t *descA, const float *angA, int nA,
                                         const uint32_t *descB, const float *angB, int nB,
                                         float ratio, unsigned thr, int check_orientation, int ratio_is_double,
                                         int *matches) {
    return match_bruteforce(descA, angA, nA, descB, angB, nB, ratio, thr, check_orientation != 0,
                            ratio_is_double != 0, matches);
}
extern "C" unsigned orc_match_bow(const uint32_t *descA, const float *angA, const int *nodeA, const unsigned char *eligA, int nA,
                                  const uint32_t *descB, const float *angB, const int *nodeB, const unsigned char *eligB, int nB,
                                  float ratio, unsigned thr, int check_orientation, int ratio_is_double, int *matches) {
    return match_bow(descA, angA, nodeA, eligA, nA, descB, angB, nodeB, eligB, nB, ratio, thr, check_orientation != 0,
                     ratio_is_double != 0, matches);
}
extern "C" unsigned orc_match_triangulation(const uint32_t *dA, const float *aA, const int *octA, const double *bearA, const int *nodeA,
                                            const unsigned char *eligA, int nA, const uint32_t *dB, const float *aB,
                                            const double *bearB, const int *nodeB, const unsigned char *eligB, int nB,
                                            const double *E, const float *scale_factors, float residual_deg_thr, unsigned thr,
                                            int check_orientation, int *matches) {
    return match_triangulation(dA, aA, octA, bearA, nodeA, eligA, nA, dB, aB, bearB, nodeB, eligB, nB, E, scale_factors,
                               residual_deg_thr, thr, check_orientation != 0, matches);
}
extern "C" int orc_check_epipolar(const double *b1, const double *b2, const double *E, float scale, float residual_deg_thr) {
    return check_epipolar(b1, b2, E, scale, residual_deg_thr) ? 1 : 0;
}
extern "C" int orc_angle_bin(float delta) { return AngleChecker::bin_of(delta, 1.0f / 30); }
extern "C" int orc_angle_invalid(const float *deltas, const int *ids, int n, int *invalid_out) {
    AngleChecker c;
    for (int i = 0; i < n; ++i) c.append(deltas[i], ids[i]);
    const auto inv = c.invalid();
    for (size_t i = 0; i < inv.size(); ++i) invalid_out[i] = inv[i];
    return (int)inv.size();
}
extern "C" void orc_bin_order(const unsigned *sizes30, unsigned *order30) {
    std::vector<unsigned> idx(30);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [sizes30](const unsigned a, const unsigned b) { return sizes30[a] > sizes30[b]; });
    for (int i = 0; i < 30; ++i) order30[i] = idx[i];
}
// std::partial_sort(first, last, last) with the same comparator: what std::sort degenerates to when the
// introsort depth limit is exhausted (checks the heap branch of the GPU library's restated sort).
extern "C" void orc_bin_order_heap(const unsigned *sizes30, unsigned *order30) {
    std::vector<unsigned> idx(30);
    std::iota(idx.begin(), idx.end(), 0);
    std::partial_sort(idx.begin(), idx.end(), idx.end(),
                      [sizes30](const unsigned a, const unsigned b) { return sizes30[a] > sizes30[b]; });
    for (int i = 0; i < 30; ++i) order30[i] = idx[i];
}
