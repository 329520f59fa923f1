// ORACLE (test infrastructure only): FAST-9/16 with non-max suppression as cv::FAST computes it,
// the per-level cell grid with the ini/min threshold fallback, and the quadtree keypoint
// distribution.
//
// The reference does NOT contain this stage: feature_detector.cpp:89-98 hands each pyramid level to
// `tracker::FeatureDetector` of the parent project (absent), and only fixes the contract around it:
// per-level budget (static_settings.cpp:39-60), 19-px border filter with std::round
// (feature_detector.cpp:103-123), float level coordinates / angle 0 / octave = level (:124-131).
// The north star names "FAST corner detection with the quadtree keypoint distribution", i.e. the
// upstream OpenVSLAM orb_extractor scheme that orb_extractor.cpp:36-38 says the reference is based
// on.  This file restates that published scheme; where upstream is implementation-defined (sort of
// (count, node pointer) pairs) the tie-break is fixed here and documented in DESIGN.md.  PARITY
// UNPINNED by the reference: this oracle is the specification of the stage.
//
// cv::FAST arithmetic (OpenCV features2d, fast.cpp / fast_score.cpp, TYPE_9_16) is pinned against
// cv2.FastFeatureDetector 4.13.0 in tests/test_oracle_cv2.py.
#include "common.h"
#include <list>

namespace orc {

// Bresenham ring of radius 3, clockwise from (0,3) as OpenCV's makeOffsets orders it.
static const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline int fast_response(const uint8_t *c, int stride) {
    int d[25];
    const int v = c[0];
    for (int k = 0; k < 16; ++k) d[k] = v - c[RING_DY[k] * stride + RING_DX[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));  // dark arc: min(v-p); bright arc: min(p-v)
    }
    return best - 1;  // cornerScore: the largest threshold at which the pixel is still a corner
}

std::vector<Kp> cv_fast(const uint8_t *img, int w, int h, int stride, int thr) {
    thr = std::min(std::max(thr, 0), 255);
    std::vector<Kp> out;
    if (w < 7 || h < 7) return out;
    // Score map, 0 where not evaluated or not a corner: cv::FAST zero-fills its three score rows and
    // never evaluates the outer 3-px frame, so NMS sees zeros there.
    const int pw = w + 2;
    std::vector<uint8_t> score((size_t)pw * (h + 2), 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int r = fast_response(img + (size_t)y * stride + x, stride);
            if (r + 1 > thr) score[(size_t)(y + 1) * pw + x + 1] = (uint8_t)r;  // corner <=> s' > thr
        }
    // strict '>' against the 8 neighbours; a score of 0 (non-corner, or s' == 1 at thr 0) never wins
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t *s = &score[(size_t)(y + 1) * pw + x + 1];
            const int v = s[0];
            if (v > s[-1] && v > s[1] && v > s[-pw - 1] && v > s[-pw] && v > s[-pw + 1]
                && v > s[pw - 1] && v > s[pw] && v > s[pw + 1])
                out.push_back({x, y, v});
        }
    return out;
}

// ---- quadtree distribution (upstream OpenVSLAM distribute_keypoints_via_tree) -------------------
namespace {
struct Node {
    int bx, by, ex, ey;          // patch [begin, end)
    std::vector<int> kps;        // candidate indices, in candidate order
    bool leaf = false;           // single keypoint: never divided again
    long seq = 0;                // creation sequence number (tie-break, see below)
    std::list<Node>::iterator self;
};
using NodeList = std::list<Node>;

void divide(const Node &n, const std::vector<Kp> &c, Node child[4]) {
    const int half_x = cv_ceil((n.ex - n.bx) / 2.0);
    const int half_y = cv_ceil((n.ey - n.by) / 2.0);
    const int mx = n.bx + half_x, my = n.by + half_y;
    child[0].bx = n.bx; child[0].by = n.by; child[0].ex = mx;   child[0].ey = my;
    child[1].bx = mx;   child[1].by = n.by; child[1].ex = n.ex; child[1].ey = my;
    child[2].bx = n.bx; child[2].by = my;   child[2].ex = mx;   child[2].ey = n.ey;
    child[3].bx = mx;   child[3].by = my;   child[3].ex = n.ex; child[3].ey = n.ey;
    for (int i : n.kps) {
        int q = 0;
        if (mx <= c[i].x) q += 1;
        if (my <= c[i].y) q += 2;
        child[q].kps.push_back(i);
    }
}

// children are pushed to the FRONT of the list in order 0..3; those with > 1 keypoints join the pool
void assign_children(Node child[4], NodeList &nodes, std::vector<NodeList::iterator> &pool, long &seq) {
    for (int q = 0; q < 4; ++q) {
        if (child[q].kps.empty()) continue;
        child[q].leaf = child[q].kps.size() == 1;
        child[q].seq = seq++;
        nodes.push_front(std::move(child[q]));
        nodes.front().self = nodes.begin();
        if (!nodes.front().leaf) pool.push_back(nodes.begin());
    }
}
}  // namespace

std::vector<int> distribute(const std::vector<Kp> &c, int area_w, int area_h, int budget) {
    std::vector<int> result;
    if (c.empty() || area_w <= 0 || area_h <= 0) return result;
    const unsigned N = (unsigned)std::max(budget, 0);

    // initial nodes: round(aspect) patches along the longer side
    const double ratio = (double)area_w / area_h;
    int nx, ny;
    double dx, dy;
    if (ratio > 1) { nx = (int)std::round(ratio); ny = 1; dx = (double)area_w / nx; dy = area_h; }
    else           { nx = 1; ny = (int)std::round(1 / ratio); dx = area_w; dy = (double)area_h / ny; }
    NodeList nodes;
    long seq = 0;
    std::vector<NodeList::iterator> initial;
    for (int i = 0; i < nx * ny; ++i) {
        const int ix = i % nx, iy = i / nx;
        Node n;
        n.bx = (int)(dx * ix); n.by = (int)(dy * iy);
        n.ex = (int)(dx * (ix + 1)); n.ey = (int)(dy * (iy + 1));
        n.seq = seq++;
        nodes.push_back(std::move(n));
        initial.push_back(std::prev(nodes.end()));
    }
    for (int i = 0; i < (int)c.size(); ++i) {
        const unsigned ix = (unsigned)(c[i].x / dx), iy = (unsigned)(c[i].y / dy);
        initial.at(ix + iy * nx)->kps.push_back(i);
    }
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->kps.empty()) { it = nodes.erase(it); continue; }
        it->leaf = it->kps.size() == 1;
        it->self = it;
        ++it;
    }

    std::vector<NodeList::iterator> pool;
    bool filled = false;
    while (true) {  // whole rounds: every dividable node is divided
        const size_t prev = nodes.size();
        pool.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->leaf) { ++it; continue; }
            Node child[4];
            divide(*it, c, child);
            assign_children(child, nodes, pool, seq);  // push_front: not revisited in this round
            it = nodes.erase(it);
        }
        if (N <= nodes.size() || nodes.size() == prev) { filled = true; break; }
        // A further whole round adds at most 3 nodes per pooled node; if that could overshoot the
        // budget, finish with the count-ordered partial round.  (Upstream variants differ in this
        // test -- ORB-SLAM2: nodes + 3*pool > N, OpenVSLAM: nodes + pool > N, which lets a whole
        // round overshoot the per-level budget by far.  The reference passes the budget as
        // `maxTracks` (feature_detector.cpp:36), a maximum, so the bounded form is the spec here:
        // a level returns at most budget + 2 keypoints.)
        if (N < nodes.size() + 3 * pool.size()) { filled = false; break; }
    }
    while (!filled) {  // partial round: most populated nodes first, stop as soon as N nodes exist
        const size_t prev = nodes.size();
        std::vector<NodeList::iterator> prev_pool;
        prev_pool.swap(pool);
        // upstream sorts (count, node*) pairs descending -- pointer order is implementation-defined;
        // fixed here as: count descending, then creation sequence descending (newest first)
        std::sort(prev_pool.begin(), prev_pool.end(), [](const NodeList::iterator &a, const NodeList::iterator &b) {
            if (a->kps.size() != b->kps.size()) return a->kps.size() > b->kps.size();
            return a->seq > b->seq;
        });
        for (auto it : prev_pool) {
            Node child[4];
            divide(*it, c, child);
            assign_children(child, nodes, pool, seq);
            nodes.erase(it);
            if (N <= nodes.size()) { filled = true; break; }
        }
        if (filled || N <= nodes.size() || nodes.size() == prev) filled = true;
    }
    // strongest response per node; the first candidate (candidate order) wins ties
    result.reserve(nodes.size());
    for (const Node &n : nodes) {
        int best = n.kps[0];
        for (size_t k = 1; k < n.kps.size(); ++k)
            if (c[n.kps[k]].resp > c[best].resp) best = n.kps[k];
        result.push_back(best);
    }
    return result;
}

std::vector<Kp> detect_level(const uint8_t *img, int w, int h, int stride, int budget,
                             int ini_thr, int min_thr, std::vector<Kp> *cands_out) {
    constexpr int overlap = 6, cell = 64;
    std::vector<Kp> cands, out;
    const int min_bx = PATCH_RADIUS, min_by = PATCH_RADIUS;
    const int max_bx = w - PATCH_RADIUS, max_by = h - PATCH_RADIUS;
    const int width = max_bx - min_bx, height = max_by - min_by;
    if (width <= 0 || height <= 0) { if (cands_out) cands_out->clear(); return out; }
    const int num_cols = width / cell + 1, num_rows = height / cell + 1;
    for (int i = 0; i < num_rows; ++i) {
        const int min_y = min_by + i * cell;
        if (max_by - overlap <= min_y) continue;
        const int max_y = std::min(min_y + cell + overlap, max_by);
        for (int j = 0; j < num_cols; ++j) {
            const int min_x = min_bx + j * cell;
            if (max_bx - overlap <= min_x) continue;
            const int max_x = std::min(min_x + cell + overlap, max_bx);
            const uint8_t *sub = img + (size_t)min_y * stride + min_x;
            std::vector<Kp> in_cell = cv_fast(sub, max_x - min_x, max_y - min_y, stride, ini_thr);
            if (in_cell.empty()) in_cell = cv_fast(sub, max_x - min_x, max_y - min_y, stride, min_thr);
            for (Kp k : in_cell) cands.push_back({k.x + j * cell, k.y + i * cell, k.resp});
        }
    }
    const std::vector<int> sel = distribute(cands, width, height, budget);
    for (int i : sel) out.push_back({cands[i].x + min_bx, cands[i].y + min_by, cands[i].resp});
    if (cands_out) cands_out->swap(cands);
    return out;
}

}  // namespace orc

using namespace orc;

extern "C" int orc_fast_response(const uint8_t *center, int stride) { return fast_response(center, stride); }

extern "C" int orc_cv_fast(const uint8_t *img, int w, int h, int stride, int thr,
                           int *xs, int *ys, int *resp, int cap) {
    const std::vector<Kp> k = cv_fast(img, w, h, stride, thr);
    for (int i = 0; i < (int)k.size() && i < cap; ++i) {
        if (xs) xs[i] = k[i].x;
        if (ys) ys[i] = k[i].y;
        if (resp) resp[i] = k[i].resp;
    }
    return (int)k.size();
}

extern "C" int orc_detect_level(const uint8_t *img, int w, int h, int stride, int budget,
                                int ini_thr, int min_thr, int *xs, int *ys, int *resp, int cap,
                                int *cand_x, int *cand_y, int *cand_resp, int cand_cap, int *n_cand) {
    std::vector<Kp> cands;
    const std::vector<Kp> k = detect_level(img, w, h, stride, budget, ini_thr, min_thr, &cands);
    for (int i = 0; i < (int)k.size() && i < cap; ++i) {
        if (xs) xs[i] = k[i].x;
        if (ys) ys[i] = k[i].y;
        if (resp) resp[i] = k[i].resp;
    }
    for (int i = 0; i < (int)cands.size() && i < cand_cap; ++i) {
        if (cand_x) cand_x[i] = cands[i].x;
        if (cand_y) cand_y[i] = cands[i].y;
        if (cand_resp) cand_resp[i] = cands[i].resp;
    }
    if (n_cand) *n_cand = (int)cands.size();
    return (int)k.size();
}

extern "C" int orc_distribute(const int *cx, const int *cy, const int *cresp, int n,
                              int area_w, int area_h, int budget, int *out_idx, int cap) {
    std::vector<Kp> c(n);
    for (int i = 0; i < n; ++i) c[i] = {cx[i], cy[i], cresp[i]};
    const std::vector<int> sel = distribute(c, area_w, area_h, budget);
    for (int i = 0; i < (int)sel.size() && i < cap; ++i) out_idx[i] = sel[i];
    return (int)sel.size();
}
