// Keypoint detection: FAST-9/16 in 64-px cells with the ini -> min threshold fallback, cell-local
// non-max suppression, and the quadtree distribution of the survivors to the per-level budget.
//
// The reference does not contain this stage (feature_detector.cpp:89-98 hands every level to the
// parent project's tracker::FeatureDetector); it fixes the contract: per-level budget
// (static_settings.cpp:39-60), 19-px border (feature_detector.cpp:103-123), float level
// coordinates, angle 0, octave = level (:124-131).  The algorithm is the upstream OpenVSLAM
// scheme the north star names: cv::FAST(cell, 20, nms) else cv::FAST(cell, 7, nms) per 64-px cell
// (70-px window, of which cv::FAST evaluates the inner 64), then distribute_keypoints_via_tree.
//
// GPU formulation
//   fast_cells_kernel: one CTA per (cell, frame).  The evaluated interiors of the cells tile
//     [22, w-22) x [22, h-22) exactly, and cv::FAST's NMS never looks across a cell edge (its score
//     rows are zero outside the evaluated window), so a cell is self-contained: stage 70x70 px in
//     shared memory, compute the corner score r (the largest t with "corner at threshold t") where
//     r >= min_thr, NMS inside the cell, keep {r >= ini} if that set is non-empty else {r >= min}.
//     A pixel that is a strict local maximum of the min-threshold score map is also one of the
//     ini-threshold map (suppressed neighbours are smaller), so one score pass serves both.
//   distribute_kernel: one CTA per (level, frame).  The sequential list algorithm is restated as
//     level-synchronous rounds over an ordered node table in shared memory: every round divides
//     a prefix of a processing order (list order in whole rounds, (count desc, list position asc)
//     in the final count-ordered rounds), children go to the front of the list in reverse creation
//     order, survivors keep their relative order.  Quadrant counts use shared-memory atomics,
//     positions come from block-wide scans, so the final node list (and therefore the output
//     order of the keypoints) is identical to the sequential one.
#include "ctx.h"
#include "tma.cuh"

namespace sg {

constexpr int FAST_THREADS = 256;
constexpr int TP = 80;                 // shared tile pitch: 3 px alignment slack + 70 px window, padded to words
constexpr int TILE_ROWS = CELL + 6;    // 70
constexpr int RPW = CELL + 8;          // response map pitch: 4-byte left pad (word-aligned rows) + 64 + right pad
constexpr int TILE_SHIFT = 3;          // the window origin 19 + 64*j is always 3 past a word boundary
static_assert((EVAL_ORIGIN - FAST_BORDER) % 4 == TILE_SHIFT && CELL % 4 == 0, "tile alignment");
constexpr int MAX_SURVIVORS = CELL * CELL;

// ---- FAST-9/16 in s16x2 lanes -------------------------------------------------------------------------
// Ring: Bresenham circle of radius 3, clockwise from (0, 3) -- the order cv::FAST uses.  Differences are
// kept biased, e = 256 + v - p (1..511), two pixels per 32-bit register, so a plain 32-bit subtract
// never borrows across lanes and unsigned 16x2 min/max (native VIMNMX.U16x2 / VIMNMX3) apply.
//   corner at threshold t  <=>  some arc of 9 has all e > 256 + t (dark) or all e < 256 - t (bright)
//   cornerScore            ==   max over arcs of max(min9(e) - 256, 256 - max9(e)) - 1
// Filter (necessary condition, cheap): every arc of 9 contains one pixel of each of the 8 antipodal
// pairs, so  min_k max(e_k, e_k+8) > 256 + t  or  max_k min(e_k, e_k+8) < 256 - t  must hold.
// NOTE: the scalar formulation `max(min9, -max9)` on int is miscompiled by ptxas 12.9 (-O1 and above,
// sm_100a); the biased unsigned form below avoids the pattern.  Parity is pinned by the GPU tests
// (candidate positions and responses, bit for bit against the oracle).
__device__ __forceinline__ unsigned lane_pair(unsigned w, int which) {   // bytes (0,1) or (2,3) -> u16x2
    return which == 0 ? __byte_perm(w, 0u, 0x4140) : __byte_perm(w, 0u, 0x4342);
}

// Per-lane a > b for u16x2 lanes below 32768: bit 15 / 31 of the result (no borrow across lanes).
__device__ __forceinline__ unsigned gt16x2(unsigned a, unsigned b) {
    return ~((b | 0x80008000u) - a) & 0x80008000u;
}

// Exact score of two (unrelated) pixels at once: lane lo = pixel at ca, lane hi = pixel at cb.
template <int P>
__device__ __forceinline__ unsigned fast_score_2px(const uint8_t *ca, const uint8_t *cb) {
    const int off[16] = {3 * P, 3 * P + 1, 2 * P + 2, P + 3, 3, -P + 3, -2 * P + 2, -3 * P + 1,
                         -3 * P, -3 * P - 1, -2 * P - 2, -P - 3, -3, P - 3, 2 * P - 2, 3 * P - 1};
    const unsigned vb = ((unsigned)ca[0] | ((unsigned)cb[0] << 16)) | 0x01000100u;
    unsigned e[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) e[k] = vb - ((unsigned)ca[off[k]] | ((unsigned)cb[off[k]] << 16));
    // sliding min / max over windows of 9 of the circular sequence, by doubling: 2, 4, 8, then +1
    unsigned mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = __vminu2(e[k], e[(k + 1) & 15]); mx2[k] = __vmaxu2(e[k], e[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = __vminu2(mn2[k], mn2[(k + 2) & 15]); mx4[k] = __vmaxu2(mx2[k], mx2[(k + 2) & 15]); }
    unsigned lo = 0u, hi = 0x02000200u;    // running max of min9, running min of max9
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        lo = __vmaxu2(lo, __vminu2(__vminu2(mn4[k], mn4[(k + 4) & 15]), e[(k + 8) & 15]));
        hi = __vminu2(hi, __vmaxu2(__vmaxu2(mx4[k], mx4[(k + 4) & 15]), e[(k + 8) & 15]));
    }
    // score + 256 = max(lo, 512 - hi) - 1   (all quantities stay inside 0..511 per lane)
    return __vmaxu2(lo, 0x02000200u - hi) - 0x00010001u;
}

struct FastMaps { CUtensorMap m[SG_MAX_LEVELS]; };   // 80 x 70 box over every pyramid level

__global__ void __launch_bounds__(FAST_THREADS)
fast_cells_kernel(const __grid_constant__ GeomDev g, const __grid_constant__ FastMaps maps, int total_cells,
                  unsigned long long *cand, int *cand_count, int *err) {
    __shared__ __align__(128) uint8_t tile[TILE_ROWS * TP];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) uint8_t resp[(CELL + 2) * RPW];
    __shared__ unsigned short surv[MAX_SURVIVORS];        // y << 6 | x of the pixels passing the filter
    __shared__ unsigned short keep[(CELL / 2) * (CELL / 2)];   // NMS winners (at most one per 2x2 block)
    __shared__ unsigned short quads[(CELL / 4) * CELL];       // y << 4 | q of the quads passing the compass test
    __shared__ int s_nsurv, s_nkeep, s_nquad, s_base;

    const int tid = threadIdx.x, f = blockIdx.y + g.frame0;
    // which level / cell
    int cell = blockIdx.x, l = 0;
    for (; l < g.levels; ++l) {
        const int n = g.lv[l].cells_x * g.lv[l].cells_y;
        if (cell < n) break;
        cell -= n;
    }
    if (l >= g.levels) return;
    const LevelDev &L = g.lv[l];
    const int ci = cell / L.cells_x, cj = cell - ci * L.cells_x;
    const int ex0 = EVAL_ORIGIN + CELL * cj, ey0 = EVAL_ORIGIN + CELL * ci;   // first evaluated pixel
    const int cw = min(CELL, L.w - EVAL_ORIGIN - ex0), ch = min(CELL, L.h - EVAL_ORIGIN - ey0);
    const int ax0 = ex0 - 3 - TILE_SHIFT, wy0 = ey0 - 3;   // word-aligned window origin

    // ---- stage the 80 x 70 window with one TMA box load (zero outside the plane) -------------------------
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, TILE_ROWS * TP);
        tma_load_3d(tile, &maps.m[l], ax0, wy0, f, &s_bar);
    }
    __syncthreads();
    mbar_wait(&s_bar, 0);

    // Two passes at most: threshold ini first; only a cell that yields nothing is redone at min.
    // A pixel with score >= t is kept by NMS iff it beats its 8 neighbours' scores; neighbours below
    // t can never beat it, so a response map holding only the scores >= t gives the exact result.
    int nkeep = 0, t = g.ini_thr;
    for (int pass = 0; pass < 2; ++pass, t = g.min_thr) {
        __syncthreads();   // the previous pass has read its counters
        if (tid == 0) { s_nsurv = 0; s_nkeep = 0; s_nquad = 0; }
        for (int i = tid; i < (CELL + 2) * RPW / 4; i += FAST_THREADS) reinterpret_cast<uint32_t *>(resp)[i] = 0;
        __syncthreads();

        // ---- compass test, 4 pixels (two s16x2 pairs) per step: an arc of 9 contains two adjacent compass
        // points (k = 0, 4, 8, 12), so two adjacent ones must both be darker than v - t or both brighter
        // than v + t.  Passing quads are compacted so that the full filter runs with full warps.
        const unsigned thi = (unsigned)(256 + t) * 0x00010001u, tlo = (unsigned)(256 - t) * 0x00010001u;
        for (int i = tid; i < (CELL / 4) * ch; i += FAST_THREADS) {
            const int y = i >> 4, q = i & 15;
            if (4 * q >= cw) continue;
            // row pointers as words; pixel x = 4q sits at tile column 4q + 6, i.e. byte 2 of word q + 1
            const uint32_t *r0 = reinterpret_cast<const uint32_t *>(tile + (y + 3) * TP) + q;
            const uint32_t c0 = r0[0], c1 = r0[1], c2 = r0[2], c3 = r0[3];
            const uint32_t ctr = __funnelshift_r(c1, c2, 16);
            const uint32_t w4 = __funnelshift_r(c2, c3, 8), w12 = __funnelshift_r(c0, c1, 24);
            const uint32_t w0 = __funnelshift_r(r0[3 * (TP / 4) + 1], r0[3 * (TP / 4) + 2], 16);
            const uint32_t w8 = __funnelshift_r(r0[-3 * (TP / 4) + 1], r0[-3 * (TP / 4) + 2], 16);
            unsigned hit = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned vb = lane_pair(ctr, h) | 0x01000100u;
                const unsigned e0 = vb - lane_pair(w0, h), e4 = vb - lane_pair(w4, h);
                const unsigned e8 = vb - lane_pair(w8, h), e12 = vb - lane_pair(w12, h);
                const unsigned dk = __vmaxu2(__vmaxu2(__vminu2(e0, e4), __vminu2(e4, e8)),
                                             __vmaxu2(__vminu2(e8, e12), __vminu2(e12, e0)));
                const unsigned br = __vminu2(__vminu2(__vmaxu2(e0, e4), __vmaxu2(e4, e8)),
                                             __vminu2(__vmaxu2(e8, e12), __vmaxu2(e12, e0)));
                hit |= gt16x2(dk, thi) | gt16x2(tlo, br);    // dk > thi  or  br < tlo
            }
            if (hit) quads[atomicAdd(&s_nquad, 1)] = (unsigned short)i;
        }
        __syncthreads();

        // ---- full filter on the compacted quads: all 16 ring pixels, 8 antipodal pairs ----------------------
        const int nquad = s_nquad;
        for (int n = tid; n < nquad; n += FAST_THREADS) {
            const int i = quads[n];
            const int y = i >> 4, q = i & 15;
            const uint32_t *r0 = reinterpret_cast<const uint32_t *>(tile + (y + 3) * TP) + q;
            const uint32_t *rm3 = r0 - 3 * (TP / 4), *rp3 = r0 + 3 * (TP / 4);
            const uint32_t *rm2 = r0 - 2 * (TP / 4), *rp2 = r0 + 2 * (TP / 4);
            const uint32_t *rm1 = r0 - (TP / 4), *rp1 = r0 + (TP / 4);
            // 4-byte windows starting at column 4q+6+dx, built from words q .. q+3 of the row
            const uint32_t c0 = r0[0], c1 = r0[1], c2 = r0[2], c3 = r0[3];
            const uint32_t a1 = rp3[1], a2 = rp3[2], b1 = rm3[1], b2 = rm3[2];
            const uint32_t d0 = rp1[0], d1 = rp1[1], d2 = rp1[2], d3 = rp1[3];
            const uint32_t g0 = rm1[0], g1 = rm1[1], g2 = rm1[2], g3 = rm1[3];
            const uint32_t ctr = __funnelshift_r(c1, c2, 16);
            const uint32_t wa[8] = {
                __funnelshift_r(a1, a2, 16),   // (0, 3)   k = 0
                __funnelshift_r(a1, a2, 24),   // (1, 3)   k = 1
                rp2[2],                        // (2, 2)   k = 2
                __funnelshift_r(d2, d3, 8),    // (3, 1)   k = 3
                __funnelshift_r(c2, c3, 8),    // (3, 0)   k = 4
                __funnelshift_r(g2, g3, 8),    // (3, -1)  k = 5
                rm2[2],                        // (2, -2)  k = 6
                __funnelshift_r(b1, b2, 24)};  // (1, -3)  k = 7
            const uint32_t wb[8] = {
                __funnelshift_r(b1, b2, 16),   // (0, -3)  k = 8
                __funnelshift_r(b1, b2, 8),    // (-1, -3) k = 9
                rm2[1],                        // (-2, -2) k = 10
                __funnelshift_r(g0, g1, 24),   // (-3, -1) k = 11
                __funnelshift_r(c0, c1, 24),   // (-3, 0)  k = 12
                __funnelshift_r(d0, d1, 24),   // (-3, 1)  k = 13
                rp2[1],                        // (-2, 2)  k = 14
                __funnelshift_r(a1, a2, 8)};   // (-1, 3)  k = 15
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // min_k max(e_k, e_k+8) > 256 + t  or  max_k min(e_k, e_k+8) < 256 - t
                const unsigned vb = lane_pair(ctr, h) | 0x01000100u;
                unsigned mn = 0x02000200u, mx = 0u;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned ea = vb - lane_pair(wa[k], h), eb = vb - lane_pair(wb[k], h);
                    mn = __vminu2(mn, __vmaxu2(ea, eb));
                    mx = __vmaxu2(mx, __vminu2(ea, eb));
                }
                const unsigned hit = gt16x2(mn, thi) | gt16x2(tlo, mx);
                const int x = 4 * q + 2 * h;
                if ((hit & 0x8000u) && x < cw) surv[atomicAdd(&s_nsurv, 1)] = (unsigned short)((y << 6) | x);
                if ((hit >> 16) && x + 1 < cw) surv[atomicAdd(&s_nsurv, 1)] = (unsigned short)((y << 6) | (x + 1));
            }
        }
        __syncthreads();

        // ---- exact corner score of the survivors, two per thread -----------------------------------------
        const int nsurv = s_nsurv;
        for (int i = tid; 2 * i < nsurv; i += FAST_THREADS) {
            const int pa = surv[2 * i], pb = surv[min(2 * i + 1, nsurv - 1)];
            const uint8_t *ca = tile + ((pa >> 6) + 3) * TP + (pa & 63) + 3 + TILE_SHIFT;
            const uint8_t *cb = tile + ((pb >> 6) + 3) * TP + (pb & 63) + 3 + TILE_SHIFT;
            const unsigned sc = fast_score_2px<TP>(ca, cb);
            const int sa = (int)(sc & 0xffffu) - 256, sb = (int)(sc >> 16) - 256;
            if (sa >= t) resp[((pa >> 6) + 1) * RPW + (pa & 63) + 4] = (uint8_t)sa;
            if (sb >= t) resp[((pb >> 6) + 1) * RPW + (pb & 63) + 4] = (uint8_t)sb;
        }
        __syncthreads();

        // ---- cell-local NMS over the survivors (strict '>' against the 8 neighbours; outside = 0) --------
        for (int i = tid; i < nsurv; i += FAST_THREADS) {
            const int p = surv[i];
            const uint8_t *s = resp + ((p >> 6) + 1) * RPW + (p & 63) + 4;
            const int v = s[0];
            if (v == 0) continue;
            if (v > s[-1] && v > s[1] && v > s[-RPW - 1] && v > s[-RPW] && v > s[-RPW + 1]
                && v > s[RPW - 1] && v > s[RPW] && v > s[RPW + 1])
                keep[atomicAdd(&s_nkeep, 1)] = (unsigned short)p;
        }
        __syncthreads();
        nkeep = s_nkeep;
        if (nkeep > 0 || g.min_thr == g.ini_thr) break;
    }
    if (nkeep == 0) return;
    if (tid == 0) s_base = atomicAdd(&cand_count[f * g.levels + l], nkeep);
    __syncthreads();
    const int base = s_base;
    if (base + nkeep > L.cand_cap) { if (tid == 0) atomicExch(err, SG_ERR_OVERFLOW); return; }
    unsigned long long *out = cand + (size_t)f * g.cand_per_frame + L.cand_off + base;
    for (int i = tid; i < nkeep; i += FAST_THREADS) {
        const int p = keep[i];
        const int y = p >> 6, x = p & 63;
        const int v = resp[(y + 1) * RPW + 4 + x];
        // order key == position in the sequential candidate list: cell row, cell column, y, x
        const unsigned key = ((unsigned)ci << 22) | ((unsigned)cj << 12) | ((unsigned)y << 6) | (unsigned)x;
        out[i] = ((unsigned long long)v << 32) | key;
    }
}

// ------------------------------------------------------------------------------------------------
// Quadtree distribution
// ------------------------------------------------------------------------------------------------
constexpr int DIST_THREADS = 512;

struct NodeBox { short bx, by, ex, ey; };

// Exclusive scan of data[0..n) in place (block-wide); returns the total.  tmp: 33 ints of smem.
__device__ int block_exclusive_scan(int *data, int n, int *tmp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int b = min(tid * per, n), e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += data[i];
    // scan of the per-thread sums
    int v = sum;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    __syncthreads();   // previous users of tmp are done
    if (lane == 31) tmp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (nt >> 5) ? tmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        tmp[lane] = w;   // inclusive warp totals
        if (lane == 31) tmp[32] = w;
    }
    __syncthreads();
    int run = v - sum + (wid ? tmp[wid - 1] : 0);
    for (int i = b; i < e; ++i) { const int x = data[i]; data[i] = run; run += x; }
    const int total = tmp[(nt >> 5) - 1];
    __syncthreads();
    return total;
}

__device__ __forceinline__ void cand_xy(unsigned key, int &x, int &y) {   // working-area coordinates
    x = FAST_BORDER + CELL * ((key >> 12) & 1023) + (key & 63);
    y = FAST_BORDER + CELL * (key >> 22) + ((key >> 6) & 63);
}

__device__ __forceinline__ int quadrant(const NodeBox &b, int x, int y) {
    const int mx = b.bx + ((b.ex - b.bx + 1) >> 1);   // begin + ceil(extent / 2)
    const int my = b.by + ((b.ey - b.by + 1) >> 1);
    return (mx <= x ? 1 : 0) + (my <= y ? 2 : 0);
}

__global__ void __launch_bounds__(DIST_THREADS)
distribute_kernel(const __grid_constant__ GeomDev g, int node_cap_max, const unsigned long long *cand_all,
                  uint32_t *cand_node_all, const int *cand_count, int *kp_xy, int *kp_resp, int *kp_count,
                  int *err) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const int NC = node_cap_max;
    // shared arrays
    unsigned long long *best = reinterpret_cast<unsigned long long *>(dsm);        // [NC]
    NodeBox *box0 = reinterpret_cast<NodeBox *>(best + NC);                        // [NC]
    NodeBox *box1 = box0 + NC;                                                     // [NC]
    int *cnt0 = reinterpret_cast<int *>(box1 + NC);                                // [NC]
    int *cnt1 = cnt0 + NC;                                                         // [NC]
    int *ccnt = cnt1 + NC;                                                         // [4 NC] child counts
    int *cpos = ccnt + 4 * NC;                                                     // [4 NC] child positions
    int *ord = cpos + 4 * NC;                                                      // [NC] processing index
    int *byord = ord + NC;                                                         // [NC]
    int *ps = byord + NC;                                                          // [NC + 1] scan over order
    int *kpos = ps + NC + 1;                                                       // [NC] position of kept nodes
    __shared__ int tmp[33];
    __shared__ int s_m;

    const int tid = threadIdx.x, l = blockIdx.x, f = blockIdx.y + g.frame0;
    const LevelDev &L = g.lv[l];
    const int N = L.budget;
    int ncand = cand_count[f * g.levels + l];
    if (ncand > L.cand_cap) ncand = 0;   // overflow already flagged by the FAST kernel
    const unsigned long long *cand = cand_all + (size_t)f * g.cand_per_frame + L.cand_off;
    uint32_t *cnode = cand_node_all + (size_t)f * g.cand_per_frame + L.cand_off;
    int *out_xy = kp_xy + (size_t)f * g.det_cap + L.kp_off;
    int *out_resp = kp_resp + (size_t)f * g.det_cap + L.kp_off;
    if (ncand == 0 || L.area_w <= 0 || L.area_h <= 0) {
        if (tid == 0) kp_count[f * g.levels + l] = 0;
        return;
    }

    NodeBox *box = box0, *nbox = box1;
    int *cnt = cnt0, *ncnt = cnt1;

    // ---- initial nodes: round(aspect) patches along the longer side -----------------------------------
    const int nx = L.init_nx, ny = L.init_ny;
    const double dx = nx > 1 ? (double)L.area_w / nx : (double)L.area_w;
    const double dy = ny > 1 ? (double)L.area_h / ny : (double)L.area_h;
    int n = nx * ny;
    for (int i = tid; i < n; i += DIST_THREADS) {
        const int ix = i % nx, iy = i / nx;
        NodeBox b;
        b.bx = (short)(int)(dx * ix); b.by = (short)(int)(dy * iy);
        b.ex = (short)(int)(dx * (ix + 1)); b.ey = (short)(int)(dy * (iy + 1));
        nbox[i] = b;
        ncnt[i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < ncand; i += DIST_THREADS) {
        int x, y;
        cand_xy((unsigned)cand[i], x, y);
        const unsigned ix = (unsigned)(x / dx), iy = (unsigned)(y / dy);
        const int nd = min((int)(ix + iy * nx), n - 1);
        cnode[i] = nd;
        atomicAdd(&ncnt[nd], 1);
    }
    __syncthreads();
    // drop empty initial nodes, keeping their order
    for (int i = tid; i < n; i += DIST_THREADS) kpos[i] = ncnt[i] > 0 ? 1 : 0;
    __syncthreads();
    {
        const int kept = block_exclusive_scan(kpos, n, tmp);
        for (int i = tid; i < n; i += DIST_THREADS)
            if (ncnt[i] > 0) { box[kpos[i]] = nbox[i]; cnt[kpos[i]] = ncnt[i]; }
        __syncthreads();
        for (int i = tid; i < ncand; i += DIST_THREADS) cnode[i] = kpos[cnode[i]];
        n = kept;
        __syncthreads();
    }

    // ---- rounds ----------------------------------------------------------------------------------------
    bool partial = false;
    while (true) {
        // A/B: quadrant populations of every dividable node
        for (int i = tid; i < 4 * n; i += DIST_THREADS) ccnt[i] = 0;
        __syncthreads();
        for (int i = tid; i < ncand; i += DIST_THREADS) {
            const int nd = cnode[i];
            if (cnt[nd] > 1) {
                int x, y;
                cand_xy((unsigned)cand[i], x, y);
                atomicAdd(&ccnt[4 * nd + quadrant(box[nd], x, y)], 1);
            }
        }
        __syncthreads();
        // C: processing order of the dividable nodes
        int P;
        if (!partial) {
            for (int j = tid; j < n; j += DIST_THREADS) ord[j] = cnt[j] > 1 ? 1 : 0;
            __syncthreads();
            P = block_exclusive_scan(ord, n, tmp);
        } else {
            // most populated first; equal counts: the node nearer the list front (the newer one) first
            for (int j = tid; j < n; j += DIST_THREADS) {
                const int cj = cnt[j];
                int r = 0;
                if (cj > 1)
                    for (int k = 0; k < n; ++k) {
                        const int ck = cnt[k];
                        r += (ck > 1 && (ck > cj || (ck == cj && k < j))) ? 1 : 0;
                    }
                ord[j] = r;
            }
            for (int j = tid; j < n; j += DIST_THREADS) kpos[j] = cnt[j] > 1 ? 1 : 0;
            __syncthreads();
            P = block_exclusive_scan(kpos, n, tmp);
        }
        for (int j = tid; j < n; j += DIST_THREADS)
            if (cnt[j] > 1) byord[ord[j]] = j;
        __syncthreads();
        // D: children created before each processed node; in a partial round, where to stop
        for (int i = tid; i < P; i += DIST_THREADS) {
            const int *c = ccnt + 4 * byord[i];
            ps[i] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
        }
        if (tid == 0) { ps[P] = 0; s_m = P; }
        __syncthreads();
        block_exclusive_scan(ps, P + 1, tmp);   // ps[i] = children created before order index i; ps[P] = all
        if (partial) {
            // list size after processing order index i:  n + (ps[i+1] - (i+1))
            for (int i = tid; i < P; i += DIST_THREADS) {
                const bool reached = N <= n + ps[i + 1] - (i + 1);
                const bool before = i > 0 && N <= n + ps[i] - i;
                if (reached && !before) s_m = i + 1;
            }
            __syncthreads();
        }
        const int m = s_m;
        const int total_new = ps[m];
        const int n_new = total_new + (n - m);
        if (n_new > NC) { if (tid == 0) { atomicExch(err, SG_ERR_OVERFLOW); kp_count[f * g.levels + l] = 0; } return; }
        // E: survivors keep their relative order behind the new children
        for (int j = tid; j < n; j += DIST_THREADS) kpos[j] = (cnt[j] > 1 && ord[j] < m) ? 0 : 1;
        __syncthreads();
        block_exclusive_scan(kpos, n, tmp);
        // F: new node table
        for (int j = tid; j < n; j += DIST_THREADS) {
            if (cnt[j] > 1 && ord[j] < m) {
                const NodeBox b = box[j];
                const int mx = b.bx + ((b.ex - b.bx + 1) >> 1), my = b.by + ((b.ey - b.by + 1) >> 1);
                int created = ps[ord[j]];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = ccnt[4 * j + q];
                    if (c == 0) continue;
                    const int pos = total_new - 1 - created++;   // pushed to the front in creation order
                    NodeBox nb;
                    nb.bx = (q & 1) ? (short)mx : b.bx; nb.ex = (q & 1) ? b.ex : (short)mx;
                    nb.by = (q & 2) ? (short)my : b.by; nb.ey = (q & 2) ? b.ey : (short)my;
                    nbox[pos] = nb;
                    ncnt[pos] = c;
                    cpos[4 * j + q] = pos;
                }
            } else {
                const int pos = total_new + kpos[j];
                nbox[pos] = box[j];
                ncnt[pos] = cnt[j];
                kpos[j] = pos;
            }
        }
        __syncthreads();
        // G: move the candidates
        for (int i = tid; i < ncand; i += DIST_THREADS) {
            const int nd = cnode[i];
            if (cnt[nd] > 1 && ord[nd] < m) {
                int x, y;
                cand_xy((unsigned)cand[i], x, y);
                cnode[i] = cpos[4 * nd + quadrant(box[nd], x, y)];
            } else {
                cnode[i] = kpos[nd];
            }
        }
        __syncthreads();
        // H: termination (uniform)
        const int n_old = n;
        n = n_new;
        { NodeBox *tb = box; box = nbox; nbox = tb; int *tc = cnt; cnt = ncnt; ncnt = tc; }
        if (N <= n || n == n_old) break;
        if (!partial) {
            // dividable nodes of the new list (all of them are children made in this round)
            for (int j = tid; j < n; j += DIST_THREADS) kpos[j] = cnt[j] > 1 ? 1 : 0;
            __syncthreads();
            const int pool = block_exclusive_scan(kpos, n, tmp);
            if (N < n + 3 * pool) partial = true;
        }
    }

    // ---- strongest candidate of every node; earlier candidate wins ties --------------------------------
    for (int j = tid; j < n; j += DIST_THREADS) best[j] = 0ull;
    __syncthreads();
    for (int i = tid; i < ncand; i += DIST_THREADS) {
        const unsigned long long c = cand[i];
        atomicMax(&best[cnode[i]], (c & 0xffffffff00000000ull) | (0xffffffffu - (unsigned)c));
    }
    __syncthreads();
    for (int j = tid; j < n; j += DIST_THREADS) {
        const unsigned long long b = best[j];
        int x, y;
        cand_xy(0xffffffffu - (unsigned)b, x, y);
        out_xy[j] = (x + PATCH_RADIUS) | ((y + PATCH_RADIUS) << 16);
        out_resp[j] = (int)(b >> 32);
    }
    if (tid == 0) kp_count[f * g.levels + l] = n;
}

size_t distribute_smem_bytes(int node_cap_max) {
    const size_t NC = node_cap_max;
    return NC * (8 + 2 * sizeof(NodeBox) + 2 * 4 + 4 * 4 + 4 * 4 + 4 + 4 + 4 + 4) + 16;
}

int launch_detect(sg_ctx *ctx, int n_frames) {
    GeomDev g = ctx->geom;
    g.frame0 = ctx->frame0;
    SG_CUDA(ctx, cudaMemsetAsync(ctx->d_cand_count + (size_t)g.frame0 * g.levels, 0, sizeof(int) * (size_t)n_frames * g.levels, ctx->stream));
    int total_cells = 0, nc_max = 1;
    for (int l = 0; l < g.levels; ++l) {
        total_cells += g.lv[l].cells_x * g.lv[l].cells_y;
        nc_max = std::max(nc_max, g.lv[l].node_cap);
    }
    if (total_cells > 0) {
        FastMaps maps;
        for (int l = 0; l < g.levels; ++l) maps.m[l] = ctx->lv[l].map_fast;
        fast_cells_kernel<<<dim3(total_cells, n_frames), FAST_THREADS, 0, ctx->stream>>>(
            g, maps, total_cells, ctx->d_cand, ctx->d_cand_count, ctx->d_err);
        SG_LAUNCH_CHECK(ctx);
    }
    mark(ctx, EV_FAST1);
    const size_t smem = distribute_smem_bytes(nc_max);
    if (smem > 48 * 1024)
        SG_CUDA(ctx, cudaFuncSetAttribute(distribute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    distribute_kernel<<<dim3(g.levels, n_frames), DIST_THREADS, smem, ctx->stream>>>(
        g, nc_max, ctx->d_cand, ctx->d_cand_node, ctx->d_cand_count, ctx->d_kp_xy, ctx->d_kp_resp, ctx->d_kp_count, ctx->d_err);
    SG_LAUNCH_CHECK(ctx);
    mark(ctx, EV_DIST1);
    ctx->detected = true;
    return SG_OK;
}

}  // namespace sg
