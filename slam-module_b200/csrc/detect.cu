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

namespace sg {

constexpr int FAST_THREADS = 256;
constexpr int TP = 80;                 // shared tile pitch (70 px + up to 3 px alignment slack, padded)
constexpr int TILE_ROWS = CELL + 6;    // 70
constexpr int RP = CELL + 2;           // response map pitch (1-px zero frame)

// ---- FAST-9/16 corner score, two horizontally adjacent pixels per 32-bit register (s16x2 lanes) ----
// cornerScore = the largest t for which the pixel is a FAST-9 corner: max over the 16 arcs of 9
// contiguous ring pixels of min(v - p) (dark arc) and min(p - v) (bright arc), minus one.
// Ring: Bresenham circle of radius 3, clockwise from (0, 3) -- the order cv::FAST uses.
// sm_100a has native 16x2 integer min/max (VIMNMX.S16x2, VIMNMX3.S16x2) and VABSDIFF4; the
// differences (|d| <= 255) fit the s16 lanes exactly.
// NOTE: a scalar formulation `max(min9, -max9)` is miscompiled by ptxas 12.9 at -O1 and above for
// sm_100a (wrong scores; correct with -Xptxas -O0).  This formulation avoids the pattern and is
// checked bit-for-bit against the oracle by the GPU parity tests (candidate responses).
__device__ __forceinline__ unsigned pack2(const uint8_t *c, int off) {
    return (unsigned)c[off] | ((unsigned)c[off + 1] << 16);
}

template <int P>
__device__ __forceinline__ unsigned fast_score2(const uint8_t *c, unsigned v, unsigned p0, unsigned p4,
                                                unsigned p8, unsigned p12) {
    unsigned d[16];
    d[0] = __vsub2(v, p0);                     d[1] = __vsub2(v, pack2(c, 3 * P + 1));
    d[2] = __vsub2(v, pack2(c, 2 * P + 2));    d[3] = __vsub2(v, pack2(c, P + 3));
    d[4] = __vsub2(v, p4);                     d[5] = __vsub2(v, pack2(c, -P + 3));
    d[6] = __vsub2(v, pack2(c, -2 * P + 2));   d[7] = __vsub2(v, pack2(c, -3 * P + 1));
    d[8] = __vsub2(v, p8);                     d[9] = __vsub2(v, pack2(c, -3 * P - 1));
    d[10] = __vsub2(v, pack2(c, -2 * P - 2));  d[11] = __vsub2(v, pack2(c, -P - 3));
    d[12] = __vsub2(v, p12);                   d[13] = __vsub2(v, pack2(c, P - 3));
    d[14] = __vsub2(v, pack2(c, 2 * P - 2));   d[15] = __vsub2(v, pack2(c, 3 * P - 1));
    // sliding min / max over windows of 9 of the circular sequence, by doubling: 2, 4, 8, then +1
    unsigned mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = __vmins2(d[k], d[(k + 1) & 15]); mx2[k] = __vmaxs2(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = __vmins2(mn2[k], mn2[(k + 2) & 15]); mx4[k] = __vmaxs2(mx2[k], mx2[(k + 2) & 15]); }
    unsigned best = 0xff00ff00u;   // (-256, -256)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const unsigned mn9 = __vimin3_s16x2(mn4[k], mn4[(k + 4) & 15], d[(k + 8) & 15]);
        const unsigned mx9 = __vimax3_s16x2(mx4[k], mx4[(k + 4) & 15], d[(k + 8) & 15]);
        best = __vimax3_s16x2(best, mn9, __vneg2(mx9));
    }
    return __vsub2(best, 0x00010001u);
}

__global__ void __launch_bounds__(FAST_THREADS)
fast_cells_kernel(const __grid_constant__ GeomDev g, const uint8_t *level0, int level0_pitch,
                  unsigned long long level0_stride, int total_cells, unsigned long long *cand,
                  int *cand_count, int *err) {
    __shared__ __align__(16) uint8_t tile[TILE_ROWS * TP];
    __shared__ uint8_t resp[RP * RP];
    __shared__ int s_cnt_ini, s_cnt_min, s_base, s_emit;

    const int tid = threadIdx.x, f = blockIdx.y;
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
    const uint8_t *img = l == 0 ? level0 + (size_t)f * level0_stride : L.pyr + (size_t)f * L.frame_stride;
    const int pitch = l == 0 ? level0_pitch : L.pitch;
    const int ex0 = EVAL_ORIGIN + CELL * cj, ey0 = EVAL_ORIGIN + CELL * ci;   // first evaluated pixel
    const int cw = min(CELL, L.w - EVAL_ORIGIN - ex0), ch = min(CELL, L.h - EVAL_ORIGIN - ey0);
    const int wx0 = ex0 - 3, wy0 = ey0 - 3;         // window origin
    const int ax0 = wx0 & ~3, shift = wx0 - ax0;    // aligned load origin

    if (tid == 0) { s_cnt_ini = 0; s_cnt_min = 0; s_emit = 0; }
    for (int i = tid; i < RP * RP / 4; i += FAST_THREADS) reinterpret_cast<uint32_t *>(resp)[i] = 0;
    // ---- stage the (cw+6) x (ch+6) window ------------------------------------------------------------
    {
        const int nwords = (shift + cw + 6 + 3) >> 2;
        for (int i = tid; i < nwords * (ch + 6); i += FAST_THREADS) {
            const int r = i / nwords, wd = i - r * nwords;
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(img + (size_t)(wy0 + r) * pitch + ax0 + 4 * wd));
            *reinterpret_cast<uint32_t *>(tile + r * TP + 4 * wd) = v;
        }
    }
    __syncthreads();

    // ---- corner score of every evaluated pixel (only where it reaches min_thr) ------------------------
    // two pixels per step; a pair is skipped when both fail the antipodal quick test: any arc of 9
    // contains one pixel of each antipodal pair, so both of a pair inside [v-t, v+t] => no corner
    const int t = g.min_thr;
    for (int i = tid; i < (CELL / 2) * ch; i += FAST_THREADS) {
        const int y = i >> 5, x = (i & 31) * 2;
        if (x >= cw) continue;
        const uint8_t *c = tile + (y + 3) * TP + x + 3 + shift;
        const unsigned v = pack2(c, 0);
        const unsigned p0 = pack2(c, 3 * TP), p8 = pack2(c, -3 * TP), p4 = pack2(c, 3), p12 = pack2(c, -3);
        const unsigned m = __vminu2(__vmaxu2(__vabsdiffu4(v, p0), __vabsdiffu4(v, p8)),
                                    __vmaxu2(__vabsdiffu4(v, p4), __vabsdiffu4(v, p12)));
        if ((int)(m & 0xffffu) <= t && (int)(m >> 16) <= t) continue;
        const unsigned sc = fast_score2<TP>(c, v, p0, p4, p8, p12);
        const int s0 = (short)(sc & 0xffffu), s1 = (short)(sc >> 16);
        if (s0 >= t) resp[(y + 1) * RP + x + 1] = (uint8_t)s0;
        if (s1 >= t && x + 1 < cw) resp[(y + 1) * RP + x + 2] = (uint8_t)s1;
    }
    __syncthreads();

    // ---- cell-local NMS (strict '>' against the 8 neighbours; outside the cell counts as 0) -----------
    unsigned keep = 0;   // bit k: pixel (tid + k*256) is a local maximum
    int n_ini = 0, n_min = 0;
    for (int k = 0; k * FAST_THREADS < CELL * ch; ++k) {
        const int i = tid + k * FAST_THREADS;
        const int y = i >> 6, x = i & 63;
        if (y >= ch || x >= cw) continue;
        const uint8_t *s = resp + (y + 1) * RP + x + 1;
        const int v = s[0];
        if (v == 0) continue;
        if (v > s[-1] && v > s[1] && v > s[-RP - 1] && v > s[-RP] && v > s[-RP + 1]
            && v > s[RP - 1] && v > s[RP] && v > s[RP + 1]) {
            keep |= 1u << k;
            ++n_min;
            if (v >= g.ini_thr) ++n_ini;
        }
    }
    n_ini = __reduce_add_sync(0xffffffffu, n_ini);
    n_min = __reduce_add_sync(0xffffffffu, n_min);
    if ((tid & 31) == 0 && n_min) { atomicAdd(&s_cnt_ini, n_ini); atomicAdd(&s_cnt_min, n_min); }
    __syncthreads();
    const bool use_ini = s_cnt_ini > 0;
    const int n_emit = use_ini ? s_cnt_ini : s_cnt_min;
    if (n_emit == 0) return;
    if (tid == 0) s_base = atomicAdd(&cand_count[f * g.levels + l], n_emit);
    __syncthreads();
    const int base = s_base;
    if (base + n_emit > L.cand_cap) { if (tid == 0) atomicExch(err, SG_ERR_OVERFLOW); return; }
    unsigned long long *out = cand + (size_t)f * g.cand_per_frame + L.cand_off + base;
    const int thr_cell = use_ini ? g.ini_thr : 1;
    while (keep) {
        const int k = __ffs(keep) - 1;
        keep &= keep - 1;
        const int i = tid + k * FAST_THREADS;
        const int y = i >> 6, x = i & 63;
        const int v = resp[(y + 1) * RP + x + 1];
        if (v >= thr_cell) {
            const int slot = atomicAdd(&s_emit, 1);
            // order key == position in the sequential candidate list: cell row, cell column, y, x
            const unsigned key = ((unsigned)ci << 22) | ((unsigned)cj << 12) | ((unsigned)y << 6) | (unsigned)x;
            out[slot] = ((unsigned long long)v << 32) | key;
        }
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

    const int tid = threadIdx.x, l = blockIdx.x, f = blockIdx.y;
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
    const GeomDev &g = ctx->geom;
    SG_CUDA(ctx, cudaMemsetAsync(ctx->d_cand_count, 0, sizeof(int) * (size_t)n_frames * g.levels, ctx->stream));
    int total_cells = 0, nc_max = 1;
    for (int l = 0; l < g.levels; ++l) {
        total_cells += g.lv[l].cells_x * g.lv[l].cells_y;
        nc_max = std::max(nc_max, g.lv[l].node_cap);
    }
    if (total_cells > 0) {
        fast_cells_kernel<<<dim3(total_cells, n_frames), FAST_THREADS, 0, ctx->stream>>>(
            g, ctx->level0, ctx->level0_pitch, ctx->level0_stride, total_cells, ctx->d_cand, ctx->d_cand_count, ctx->d_err);
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
