// CPU oracle (test infrastructure only) for the "next" rows of SURVEY.md 8(f):
//   MapPoint::updateDescriptor           map_point.cpp:75-116      (Hamming medoid of a map point's observations)
//   FeatureSearchImplementation          feature_search.cpp:22-48  (Y-sorted index, radius query)
//   candidate loops of searchByProjection keyframe_matcher.cpp:356-386, replaceDuplication :482-499 and
//   findMatchesTranformedMps :604-627   (best / second-best Hamming over the radius query result)
// The geometry in front of those loops (reprojection, viewing distance, scale prediction) stays with the
// caller: the oracle and the CUDA library take the projected point, the radius and the descriptor.
#include <algorithm>
#include <vector>
#include "common.h"

extern "C" unsigned orc_hamming(const uint32_t *a, const uint32_t *b);   // match.cpp (openvslam/match_base.h:18-39)

namespace orc {
static inline unsigned hamming(const uint32_t *a, const uint32_t *b) { return orc_hamming(a, b); }

// map_point.cpp:87-115.  Returns best_idx (0 for an empty / degenerate input, like the reference's default).
static int medoid(const uint32_t *desc, int n) {
    if (n <= 0) return 0;
    std::vector<std::vector<unsigned>> d(n, std::vector<unsigned>(n, 0));
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) d[i][j] = d[j][i] = hamming(desc + 8 * i, desc + 8 * j);
    unsigned best_median = 256;
    int best = 0;
    for (int i = 0; i < n; ++i) {
        std::vector<unsigned> row(d[i]);
        std::sort(row.begin(), row.end());
        const unsigned med = row[static_cast<unsigned>(0.5 * (n - 1))];
        if (med < best_median) { best_median = med; best = i; }
    }
    return best;
}

struct Node { float x, y; int idx; };

// feature_search.cpp:22-31: std::sort by y (the order of equal y is libstdc++'s; the library under test is
// given this very order, as the reference's matchers are)
static std::vector<Node> build_index(const float *x, const float *y, int n) {
    std::vector<Node> v(n);
    for (int i = 0; i < n; ++i) v[i] = Node{x[i], y[i], i};
    std::sort(v.begin(), v.end(), [](const Node &a, const Node &b) { return a.y < b.y; });
    return v;
}

// feature_search.cpp:33-48
static void around(const std::vector<Node> &index, float x, float y, float r, std::vector<int> &out) {
    out.clear();
    const Node lb{x, y - r, 0};
    for (auto it = std::lower_bound(index.begin(), index.end(), lb, [](const Node &a, const Node &b) { return a.y < b.y; });
         it != index.end() && it->y <= y + r; ++it) {
        const float dx = x - it->x, dy = y - it->y;
        if (dx * dx + dy * dy < r * r) out.push_back(it->idx);
    }
}
}  // namespace orc

using namespace orc;

extern "C" void orc_medoid(const uint32_t *desc, const long long *offsets, int n_seg, int *best) {
    for (int s = 0; s < n_seg; ++s) best[s] = medoid(desc + 8 * offsets[s], (int)(offsets[s + 1] - offsets[s]));
}

extern "C" void orc_feature_index(const float *x, const float *y, int n, int *order) {
    const auto idx = build_index(x, y, n);
    for (int i = 0; i < n; ++i) order[i] = idx[i].idx;
}

extern "C" int orc_features_around(const float *x, const float *y, int n, float qx, float qy, float r, int *out) {
    const auto idx = build_index(x, y, n);
    std::vector<int> o;
    around(idx, qx, qy, r, o);
    for (size_t i = 0; i < o.size(); ++i) out[i] = o[i];
    return (int)o.size();
}

// The candidate loop shared by the projection matchers, queries processed in order.
//   mode 0: best only (replaceDuplication :482-499 with thr 50; findMatchesTranformedMps :604-627 with thr 100 and
//           the octave window [pred - 1, pred] when q_pred_level != NULL); a query never consumes a keypoint.
//   mode 1: searchByProjection :356-386: skip keypoints that are taken (`taken` in/out: initially the keypoints
//           that already own an observed map point, :358), best and second best with their levels, accept when
//           best <= thr and not (same level and best > 0.8 * second), the accepted keypoint becomes taken.
// out_idx[q] = matched keypoint or -1, out_dist[q] = its distance (256 when unmatched).
extern "C" int orc_search_candidates(const float *kx, const float *ky, const int *koct, const uint32_t *kdesc, int nK,
                                     unsigned char *taken, const float *qx, const float *qy, const float *qr,
                                     const uint32_t *qdesc, const int *q_pred_level, int nQ, int mode, unsigned thr,
                                     int *out_idx, unsigned *out_dist) {
    const auto index = build_index(kx, ky, nK);
    std::vector<int> cand;
    int count = 0;
    for (int q = 0; q < nQ; ++q) {
        out_idx[q] = -1;
        out_dist[q] = 256;
        around(index, qx[q], qy[q], qr[q], cand);
        if (cand.empty()) continue;
        if (mode == 0) {
            unsigned best = 256;
            int best_idx = -1;
            for (int i : cand) {
                if (q_pred_level && (koct[i] < q_pred_level[q] - 1 || koct[i] > q_pred_level[q])) continue;
                const unsigned d = hamming(qdesc + 8 * q, kdesc + 8 * i);
                if (d < best) { best = d; best_idx = i; }
            }
            if (best_idx != -1 && best <= thr) { out_idx[q] = best_idx; out_dist[q] = best; ++count; }
        } else {
            int best = 256, best2 = 256, lvl = -1, lvl2 = -1, best_idx = -1;
            for (int i : cand) {
                if (taken[i]) continue;
                const int d = (int)hamming(qdesc + 8 * q, kdesc + 8 * i);
                if (d < best) { best2 = best; best = d; lvl2 = lvl; lvl = koct[i]; best_idx = i; }
                else if (d < best2) { lvl2 = koct[i]; best2 = d; }
            }
            if (best_idx == -1) continue;
            if (best <= (int)thr) {
                if (lvl == lvl2 && best > 0.8 * best2) continue;
                out_idx[q] = best_idx;
                out_dist[q] = (unsigned)best;
                taken[best_idx] = 1;
                ++count;
            }
        }
    }
    return count;
}

// matchMapPointsSim3 (keyframe_matcher.cpp:633-686): findMatchesTranformedMps (:552-631) in both directions, then only
// the pairs on which the two directions agree (:672-685).  The Sim3 geometry (transform, reprojection, viewing
// distance, scale prediction) stays with the caller: q12_* is, per keypoint i of keyframe 1, the projection of its
// map point into keyframe 2 with the search radius margin * scaleFactor[level] and the predicted level; r < 0 marks
// a keypoint that issues no query (no map point, seeded as already matched :643-649, not TRIANGULATED, outside the
// image or the viewing-distance range).  q21_* likewise per keypoint of keyframe 2.  out_pairs: (i, j) in i order.
extern "C" int orc_match_sim3(const float *x1, const float *y1, const int *oct1, const uint32_t *d1, int n1,
                              const float *x2, const float *y2, const int *oct2, const uint32_t *d2, int n2,
                              const float *q12x, const float *q12y, const float *q12r, const uint32_t *q12desc, const int *q12lvl,
                              const float *q21x, const float *q21y, const float *q21r, const uint32_t *q21desc, const int *q21lvl,
                              int *out_pairs) {
    auto direction = [](const float *kx, const float *ky, const int *koct, const uint32_t *kdesc, int nK, const float *qx,
                        const float *qy, const float *qr, const uint32_t *qdesc, const int *qlvl, int nQ) {
        const auto index = build_index(kx, ky, nK);
        std::vector<int> match(nQ, -1), cand;
        for (int q = 0; q < nQ; ++q) {
            if (qr[q] < 0) continue;
            around(index, qx[q], qy[q], qr[q], cand);
            unsigned best = 256;
            int best_idx = -1;
            for (int i : cand) {
                if (koct[i] < qlvl[q] - 1 || koct[i] > qlvl[q]) continue;          // :611
                const unsigned d = hamming(qdesc + 8 * q, kdesc + 8 * i);
                if (d < best) { best = d; best_idx = i; }
            }
            if (best <= 100) match[q] = best_idx;                                   // :625 HAMMING_DIST_THR_HIGH
        }
        return match;
    };
    const std::vector<int> m12 = direction(x2, y2, oct2, d2, n2, q12x, q12y, q12r, q12desc, q12lvl, n1);
    const std::vector<int> m21 = direction(x1, y1, oct1, d1, n1, q21x, q21y, q21r, q21desc, q21lvl, n2);
    int n = 0;
    for (int i = 0; i < n1; ++i) {
        const int j = m12[i];
        if (j < 0) continue;
        if (m21[j] == i) { out_pairs[2 * n] = i; out_pairs[2 * n + 1] = j; ++n; }
    }
    return n;
}
