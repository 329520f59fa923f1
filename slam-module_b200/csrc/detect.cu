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
constexpr int RPW = TP;                // response map pitch: 4-byte left pad (word-aligned rows) + 64 + right pad; equal to the
                                       // window pitch, so that one queue entry addresses the pair words AND the response bytes
constexpr int TILE_SHIFT = 3;          // the window origin 19 + 64*j is always 3 past a word boundary
static_assert((EVAL_ORIGIN - FAST_BORDER) % 4 == TILE_SHIFT && CELL % 4 == 0, "tile alignment");
static_assert(((CELL + 2) * TP) % 16 == 0 && TP == 2 * (TP / 2) && (CELL + 6) * 80 % 8 == 0, "vector widths of the zeroing / pair-word loops");
constexpr int WP = TP / 2;             // pair-word pitch: 40 words per row
constexpr int MAX_ENTRIES = CELL * CELL;   // 2048 pixel pairs, each at most twice (both polarities)
constexpr int WARP_Q = MAX_ENTRIES / 2 / (FAST_THREADS / 32);   // per-warp queue: 8 rows x 32 pairs
constexpr int MAX_SCORED = 1024;           // list of scored pixels (NMS candidates); beyond it the map is scanned

// ---- FAST-9/16 in u16x2 lanes ---------------------------------------------------------------------------
// Ring: Bresenham circle of radius 3, clockwise from (0, 3) -- the order cv::FAST uses.  Two horizontally
// adjacent pixels (x, x+1) are processed per 32-bit register, one per 16-bit lane.  The cell window is kept in
// shared memory as PAIR WORDS: We[y][i] = p(2i) | p(2i+1) << 16 and Wo[y][i] = p(2i+1) | p(2i+2) << 16, so the
// ring pixel of both lanes at any offset (dx, dy) is ONE aligned 32-bit load (We for even dx, Wo for odd dx):
// no per-use byte permutes.  Differences are kept biased, e = 256 + v - p (1..511), so a plain 32-bit subtract
// never borrows across lanes and the native 16x2 min / max (VIMNMX.U16x2, VIMNMX3) apply.
//   corner at threshold t (dark)  <=>  some arc of 9 has all e > 256 + t
//   dark score                    ==   max over arcs of min9(e) - 257        (cv::cornerScore convention)
// The bright polarity is the dark polarity of the complemented image (p -> 255 - p), so a lane that needs the
// bright test XORs its pixels with 0xff and runs the same code.
// Stage 1 (every pair): an arc of 9 contains one pixel of each antipodal pair, so for the compass and the diagonal pairs
//   (e0 or e8 dark) and (e4 or e12 dark) and (e2 or e10 dark) and (e6 or e14 dark)
//   <=>  min(max(e0, e8), max(e4, e12), max(e2, e10), max(e6, e14)) > 256 + t   is necessary.
// Stage 2 (pairs that pass, compacted): exact one-sided score of both lanes for the polarity stage 1 left
// possible (a pixel can be a corner of one polarity only; a lane that passes both is queued twice).
// NOTE: the scalar formulation `max(min9, -max9)` on int is miscompiled by ptxas 12.9 (-O1 and above,
// sm_100a); the biased unsigned form avoids the pattern.  Parity is pinned by the GPU tests (candidate
// positions and responses, bit for bit against the oracle).

// Per-lane a > b for u16x2 lanes below 32768: bit 15 / 31 of the result (no borrow across lanes).
__device__ __forceinline__ unsigned gt16x2(unsigned a, unsigned b) {
    return ~((b | 0x80008000u) - a) & 0x80008000u;
}

struct FastMaps { CUtensorMap m[SG_MAX_LEVELS]; };   // 80 x 70 box over every pyramid level

__global__ void __launch_bounds__(FAST_THREADS, 5)
fast_cells_kernel(const __grid_constant__ GeomDev g, const __grid_constant__ FastMaps maps, const int4 *cells,
                  unsigned long long *cand, int *cand_count, int *err) {
    // the byte window is dead once the pair words are built: the stage-1 queues and the NMS winners reuse its bytes
    __shared__ __align__(128) uint8_t tile[MAX_ENTRIES + (CELL / 2) * (CELL / 2) * 2];
    static_assert(sizeof(tile) >= TILE_ROWS * TP + 16, "window fits under the queues");
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) uint32_t We[TILE_ROWS * WP], Wo[TILE_ROWS * WP];
    __shared__ __align__(16) uint8_t resp[(CELL + 2) * RPW];
    unsigned short *ent = reinterpret_cast<unsigned short *>(tile);   // [MAX_ENTRIES / 2] per-warp queues of pairs that pass stage 1: y * WP + k
    unsigned short *keep = ent + MAX_ENTRIES / 2;                     // [(CELL / 2)^2] NMS winners (at most one per 2x2 block)
    __shared__ unsigned short scored[MAX_SCORED];              // y * RPW + x of the pixels with a score >= t
    __shared__ int s_nscored, s_nkeep, s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, f = blockIdx.y + g.frame0;
    // which level / cell: one broadcast load of the table built at sg_create (fast_cell_table)
    const int4 ce = __ldg(cells + blockIdx.x);
    const int l = ce.x & 0xff, ci = (ce.x >> 8) & 0xfff, cj = ce.x >> 20;
    const LevelDev &L = g.lv[l];
    const int ex0 = ce.y, ey0 = ce.z;                      // first evaluated pixel
    const int cw = ce.w & 0xffff, ch = ce.w >> 16;
    const int ax0 = ex0 - 3 - TILE_SHIFT, wy0 = ey0 - 3;   // word-aligned window origin

    // ---- stage the 80 x 70 window with one TMA box load (zero outside the plane) -------------------------
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, TILE_ROWS * TP);
        tma_load_3d(tile, &maps.m[l], ax0, wy0, f, &s_bar);
        s_nscored = 0; s_nkeep = 0;
    }
    for (int i = tid; i < (CELL + 2) * RPW / 16; i += FAST_THREADS) reinterpret_cast<uint4 *>(resp)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    mbar_wait(&s_bar, 0);

    // ---- pair words: 4 per aligned tile word ----------------------------------------------------------------
    // two tile words (8 pixels) per item: 64-bit load, 128-bit stores, shared-space addresses
    {
        const uint32_t t_addr = smem_u32(tile), we_addr = smem_u32(We), wo_addr = smem_u32(Wo);
        for (int i = tid; i < TILE_ROWS * (TP / 8); i += FAST_THREADS) {
            // words 2i, 2i + 1 of the tile and the word behind them; pair indices 4i .. 4i + 3 of both tiles
            const uint2 w = lds64(t_addr + 8u * i);
            const uint32_t nx = lds32(t_addr + 8u * i + 8u);
            sts128(we_addr + 16u * i, make_uint4(__byte_perm(w.x, 0u, 0x4140), __byte_perm(w.x, 0u, 0x4342),
                                                 __byte_perm(w.y, 0u, 0x4140), __byte_perm(w.y, 0u, 0x4342)));
            sts128(wo_addr + 16u * i, make_uint4(__byte_perm(w.x, 0u, 0x4241), __byte_perm(w.x, w.y, 0x0403) & 0x00ff00ffu,
                                                 __byte_perm(w.y, 0u, 0x4241), __byte_perm(w.y, nx, 0x0403) & 0x00ff00ffu));
        }
    }
    __syncthreads();

    // Two passes at most: threshold ini first; only a cell that yields nothing is redone at min.
    // A pixel with score >= t is kept by NMS iff it beats its 8 neighbours' scores; neighbours below
    // t can never beat it, so a response map holding only the scores >= t gives the exact result.
    // (Scores written by the ini pass are a subset of the min pass's, with the same values: no re-zeroing.)
    int nkeep = 0, t = g.ini_thr;
    for (int pass = 0; pass < 2; ++pass, t = g.min_thr) {
        const unsigned thi = (unsigned)(256 + t) * 0x00010001u, tlo = (unsigned)(256 - t) * 0x00010001u;
        // ---- stage 1: compass test, one warp per row, lane = pixel pair (2k, 2k + 1) ------------------------
        // Only "does either pixel of the pair pass" is decided here (the polarity flags are recomputed for the few
        // pairs that reach stage 2).  Passing pairs go to a queue PRIVATE to the warp (ballot + prefix popc, no
        // atomics, no CTA barrier): a warp owns 8 consecutive rows and scores its own queue right after.
        unsigned short *q = ent + warp * WARP_Q;
        const uint32_t q_addr = smem_u32(q);
        int nq = 0;
        const unsigned lt = (1u << lane) - 1u;
        // pixels outside the cell never count
        const unsigned okH = (2 * lane < cw ? 0x00008000u : 0u) | (2 * lane + 1 < cw ? 0x80000000u : 0u);
        const unsigned thiH = thi | 0x80008000u, Hmtlo = 0x80008000u - tlo;
        // A warp walks down ROWS_PER_WARP consecutive rows and keeps the ring rows it has loaded in registers: the word of
        // column k in row y is w0 of row y - 3, the centre of row y and w8 of row y + 3; the words of columns k - 1 / k + 1
        // serve rows y - 2 and y + 2.  5 loads per row instead of 9 (54 per 8 rows with the fill).
        constexpr int ROWS_PER_WARP = CELL / (FAST_THREADS / 32);
        const int y0 = warp * ROWS_PER_WARP;
        if (y0 < ch) {
            const uint32_t *we = We + (y0 + 3) * WP + lane + 3, *wo = Wo + (y0 + 3) * WP + lane;
            unsigned cc[7], cl[5], cr[5];        // rows y-3..y+3 of column k; rows y-2..y+2 of columns k-1 and k+1
#pragma unroll
            for (int i = 0; i < 7; ++i) cc[i] = we[(i - 3) * WP];
#pragma unroll
            for (int i = 0; i < 5; ++i) { cl[i] = we[(i - 2) * WP - 1]; cr[i] = we[(i - 2) * WP + 1]; }
            auto test_row = [&](int j) {
                const int y = y0 + j;
                const unsigned vb = cc[3] | 0x01000100u;
                // e_i = vb - w_i (biased differences 256 + v - p_i, one per 16-bit lane).  Because vb is common,
                //   dk = min_i max(e_i, e_i+8) = vb - max_i min(w_i, w_i+8),  br = max_i min(e_i, e_i+8) = vb - min_i max(w_i, w_i+8):
                // the min / max lattice runs on the raw pair words and only two subtractions remain.
                const unsigned w0 = cc[6], w8 = cc[0], w4 = wo[j * WP + 4], w12 = wo[j * WP + 1];
                // the two diagonal antipodal pairs (2, 2) / (-2, -2) and (2, -2) / (-2, 2) halve what reaches stage 2
                const unsigned w2 = cr[4], w10 = cl[0], w6 = cr[0], w14 = cl[4];
                const unsigned A = __vimax3_u16x2(__vminu2(w0, w8), __vminu2(w4, w12), __vmaxu2(__vminu2(w2, w10), __vminu2(w6, w14)));
                const unsigned B = __vimin3_u16x2(__vmaxu2(w0, w8), __vmaxu2(w4, w12), __vminu2(__vmaxu2(w2, w10), __vmaxu2(w6, w14)));
                // lane bit 15 / 31 of (thiH - dk) = thiH - vb + A is clear iff dk > thi; of (br + H) - tlo = vb - B + (H - tlo)
                // iff br < tlo.  Every 16-bit lane of both sums stays inside [0x7f00, 0x8200]: no carry between the lanes.
                const unsigned pass = ~((thiH - vb + A) & (vb - B + Hmtlo)) & okH;
                const unsigned m = __ballot_sync(0xffffffffu, pass != 0u);
                if (pass) sts16(q_addr + 2u * (unsigned)(nq + __popc(m & lt)), (unsigned)(lane + y * WP));   // shared-space store: no generic address math per row
                nq += __popc(m);
            };
            auto slide = [&](int j) {      // the windows move down one row (register renaming after unrolling)
#pragma unroll
                for (int i = 0; i < 6; ++i) cc[i] = cc[i + 1];
#pragma unroll
                for (int i = 0; i < 4; ++i) { cl[i] = cl[i + 1]; cr[i] = cr[i + 1]; }
                cc[6] = we[(j + 4) * WP];
                cl[4] = we[(j + 3) * WP - 1];
                cr[4] = we[(j + 3) * WP + 1];
            };
            if (y0 + ROWS_PER_WARP <= ch) {                    // all rows inside the cell (every cell but the last row of cells)
#pragma unroll
                for (int j = 0; j < ROWS_PER_WARP; ++j) {
                    test_row(j);
                    if (j + 1 < ROWS_PER_WARP) slide(j);
                }
            } else {
                for (int j = 0; y0 + j < ch; ++j) {
                    // (not unrolled: the windows really move)
                    const int y = y0 + j;
                    const uint32_t *wr = We + (y + 3) * WP + lane + 3;
#pragma unroll
                    for (int i = 0; i < 7; ++i) cc[i] = wr[(i - 3) * WP];
                    cl[0] = wr[-2 * WP - 1]; cl[4] = wr[2 * WP - 1]; cr[0] = wr[-2 * WP + 1]; cr[4] = wr[2 * WP + 1];
                    const unsigned vb = cc[3] | 0x01000100u;
                    const unsigned w0 = cc[6], w8 = cc[0], w4 = wo[j * WP + 4], w12 = wo[j * WP + 1];
                    const unsigned w2 = cr[4], w10 = cl[0], w6 = cr[0], w14 = cl[4];
                    const unsigned A = __vimax3_u16x2(__vminu2(w0, w8), __vminu2(w4, w12), __vmaxu2(__vminu2(w2, w10), __vminu2(w6, w14)));
                    const unsigned B = __vimin3_u16x2(__vmaxu2(w0, w8), __vmaxu2(w4, w12), __vminu2(__vmaxu2(w2, w10), __vmaxu2(w6, w14)));
                    const unsigned pass = ~((thiH - vb + A) & (vb - B + Hmtlo)) & okH;
                    const unsigned m = __ballot_sync(0xffffffffu, pass != 0u);
                    if (pass) sts16(q_addr + 2u * (unsigned)(nq + __popc(m & lt)), (unsigned)(lane + y * WP));
                    nq += __popc(m);
                }
            }
        }
        __syncwarp();

        // ---- stage 2: exact one-sided score of both lanes of every queued pair ----------------------------
        for (int i0 = 0; i0 < nq; i0 += 32) {
            const int i = i0 + lane;
            const bool live = i < nq;
            const unsigned e = live ? q[i] : 0u;
            // e = y * WP + k: the pair words of (y, k) and the response bytes of its two pixels are a base plus e
            const uint32_t *we = We + 3 * WP + 3 + e, *wo = Wo + 3 * WP + e;
            unsigned r[16];
            const unsigned c = we[0];
            r[0] = we[3 * WP];      r[1] = wo[3 * WP + 3];  r[2] = we[2 * WP + 1];   r[3] = wo[WP + 4];
            r[4] = wo[4];           r[5] = wo[-WP + 4];     r[6] = we[-2 * WP + 1];  r[7] = wo[-3 * WP + 3];
            r[8] = we[-3 * WP];     r[9] = wo[-3 * WP + 2]; r[10] = we[-2 * WP - 1]; r[11] = wo[-WP + 1];
            r[12] = wo[1];          r[13] = wo[WP + 1];     r[14] = we[2 * WP - 1];  r[15] = wo[3 * WP + 2];
            // polarity each 16-bit lane can still have, from the compass pixels: bit 15 / 31 of `dark` and `bright`
            // (dead entries and pixels outside the cell: neither)
            unsigned dark, bright;
            {
                const unsigned vb = c | 0x01000100u;
                const unsigned A = __vmaxu2(__vminu2(r[0], r[8]), __vminu2(r[4], r[12]));
                const unsigned B = __vminu2(__vmaxu2(r[0], r[8]), __vmaxu2(r[4], r[12]));
                unsigned ok = live ? 0x80008000u : 0u;
                if (cw < CELL) {                                   // last cell column: k = e mod WP (e / 40 by multiply-shift, e < 2800)
                    const int k = (int)(e - WP * ((e * 3277u) >> 17));
                    ok &= (2 * k < cw ? 0x00008000u : 0u) | (2 * k + 1 < cw ? 0x80000000u : 0u);
                }
                dark = ~(thiH - vb + A) & ok;
                bright = ~(vb - B + Hmtlo) & ok;
            }
            uint8_t *rp = resp + RPW + 4 + 2 * e;
            // dark first; a lane that can have both polarities (rare) is scored a second time as bright
            unsigned act = dark | bright, br = bright & ~dark;     // lanes scored in this round; of those, the bright ones
            const unsigned won_bias = (unsigned)(0x8000 - 257 - t) * 0x00010001u;
            for (int round = 0; round < 2; ++round) {
                const unsigned cm = (br >> 15) * 0xffu;             // 0x000000ff / 0x00ff0000: complement -> bright test
                const unsigned vb = (c ^ cm) | 0x01000100u;
                // score = max over the 16 arcs of min over the arc's 9 ring pixels of e_j = vb - r'_j
                //       = vb - min over arcs of max over the arc of r'_j   (r' = ring pixels, complemented for bright).
                // Sliding max over windows of 9 of the circular sequence as 3 x 3 with the three-input VIMNMX3:
                // m3[j] = max(d[j..j+2]), window j = max(m3[j], m3[j+3], m3[j+6]); 16 + 16 operations, then 8 for the minimum
                // over the 16 windows (the doubling scheme 2, 4, 8, +1 needs 56).
                unsigned d[16], m3[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) d[j] = r[j] ^ cm;
#pragma unroll
                for (int j = 0; j < 16; ++j) m3[j] = __vimax3_u16x2(d[j], d[(j + 1) & 15], d[(j + 2) & 15]);
                unsigned hi = 0xffffffffu;
#pragma unroll
                for (int j = 0; j < 16; j += 2)
                    hi = __vimin3_u16x2(hi, __vimax3_u16x2(m3[j], m3[(j + 3) & 15], m3[(j + 6) & 15]),
                                        __vimax3_u16x2(m3[j + 1], m3[(j + 4) & 15], m3[(j + 7) & 15]));
                const unsigned lo = vb - hi;                        // score + 257 per lane, 1 .. 511
                const unsigned won = (lo + won_bias) & act;         // bit 15 / 31: scored in this round and score >= t
                if (won) {
                    const bool w0 = (won & 0x8000u) != 0u, w1 = (won >> 31) != 0u;
                    const int s0 = (int)(lo & 0xffffu) - 257, s1 = (int)(lo >> 16) - 257;
                    if (w0) rp[0] = (uint8_t)s0;
                    if (w1) rp[1] = (uint8_t)s1;
                    // pixels with a score are the only NMS candidates (a pixel scores in at most one polarity)
                    int at = atomicAdd(&s_nscored, (w0 ? 1 : 0) + (w1 ? 1 : 0));
                    if (w0 && at < MAX_SCORED) scored[at] = (unsigned short)(2 * e);
                    at += w0 ? 1 : 0;
                    if (w1 && at < MAX_SCORED) scored[at] = (unsigned short)(2 * e + 1);
                }
                act = br = dark & bright;
                if (!__any_sync(0xffffffffu, act != 0u)) break;
            }
        }
        __syncthreads();

        // ---- cell-local NMS (strict '>' against the 8 neighbours; outside the cell = 0) --------------------
        const int nscored = s_nscored;
        if (nscored <= MAX_SCORED) {
            for (int i = tid; i < nscored; i += FAST_THREADS) {
                const int p = scored[i];
                const uint8_t *sp = resp + RPW + 4 + p;
                // all eight neighbours first, then one comparison against their maximum (no short-circuit branches)
                const int v = sp[0];
                const int up = __vimax3_s32(sp[-RPW - 1], sp[-RPW], sp[-RPW + 1]), dn = __vimax3_s32(sp[RPW - 1], sp[RPW], sp[RPW + 1]);
                if (v > __vimax3_s32(up, dn, max((int)sp[-1], (int)sp[1]))) keep[atomicAdd(&s_nkeep, 1)] = (unsigned short)p;
            }
        } else {
            // more scored pixels than the list holds (very dense corners): scan the response map instead
            for (int i = tid; i < CELL * CELL; i += FAST_THREADS) {
                const int p = (i >> 6) * RPW + (i & 63);
                const uint8_t *sp = resp + RPW + 4 + p;
                const int v = sp[0];
                if (v != 0) {
                    const int up = __vimax3_s32(sp[-RPW - 1], sp[-RPW], sp[-RPW + 1]), dn = __vimax3_s32(sp[RPW - 1], sp[RPW], sp[RPW + 1]);
                    if (v > __vimax3_s32(up, dn, max((int)sp[-1], (int)sp[1]))) keep[atomicAdd(&s_nkeep, 1)] = (unsigned short)p;
                }
            }
        }
        __syncthreads();
        nkeep = s_nkeep;
        if (nkeep > 0 || g.min_thr == g.ini_thr) break;
        if (tid == 0) s_nscored = 0;
        __syncthreads();
    }
    if (nkeep == 0) return;
    if (tid == 0) s_base = atomicAdd(&cand_count[f * g.levels + l], nkeep);
    __syncthreads();
    const int base = s_base;
    if (base + nkeep > L.cand_cap) { if (tid == 0) atomicExch(err, SG_ERR_OVERFLOW); return; }
    unsigned long long *out = cand + (size_t)f * g.cand_per_frame + L.cand_off + base;
    for (int i = tid; i < nkeep; i += FAST_THREADS) {
        const int p = keep[i];
        const int y = (int)(((unsigned)p * 3277u) >> 18), x = p - y * RPW;     // p = y * 80 + x, p < 5120
        const int v = resp[RPW + 4 + p];
        // order key == position in the sequential candidate list: cell row, cell column, y, x
        const unsigned key = ((unsigned)ci << 22) | ((unsigned)cj << 12) | ((unsigned)y << 6) | (unsigned)x;
        out[i] = ((unsigned long long)v << 32) | key;
    }
}

// ------------------------------------------------------------------------------------------------
// Quadtree distribution
// ------------------------------------------------------------------------------------------------
constexpr int DIST_THREADS = 256;

struct NodeBox { short bx, by, ex, ey; };

// Exclusive scan of data[0..n) in place (block-wide); returns the total.  tmp: 2 x 8 ints of smem, used alternately
// (`flip` toggles per call), so that one barrier between the warp totals and their use and one behind the scan suffice.
__device__ int block_exclusive_scan(int *data, int n, int *tmp, int &flip) {
    const int tid = threadIdx.x, nt = DIST_THREADS;   // every caller runs DIST_THREADS threads: a shift, not a division
    const int per = (n + nt - 1) / nt;
    const int b = min(tid * per, n), e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += data[i];
    int v = sum;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    int *t = tmp + 8 * flip;
    flip ^= 1;
    if (lane == 31) t[wid] = v;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < DIST_THREADS / 32; ++w) {
        const int x = t[w];
        base += w < wid ? x : 0;
        total += x;
    }
    int run = v - sum + base;
    for (int i = b; i < e; ++i) { const int x = data[i]; data[i] = run; run += x; }
    __syncthreads();
    return total;
}

// Number of j < n with pred(j) (block-wide), one barrier per DIST_THREADS elements.
template <class F> __device__ int block_count(int n, F pred) {
    int total = 0;
    for (int j0 = 0; j0 < n; j0 += DIST_THREADS) {
        const int j = j0 + threadIdx.x;
        total += __syncthreads_count(j < n && pred(j));
    }
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

// Candidates of one (frame, level): position and current node, either cached in shared memory (levels with at most
// CAND_SMEM candidates: every level of the VGA / 720p configurations) or in the global arrays.
constexpr int CAND_SMEM = 2688;
template <bool CACHED> struct CandStore {
    const unsigned long long *cand;
    uint32_t *cnode;
    uint32_t *s_xy;
    unsigned short *s_node;
    __device__ __forceinline__ void xy(int i, int &x, int &y) const {
        if (CACHED) { const uint32_t p = s_xy[i]; x = p & 0xffff; y = p >> 16; }
        else cand_xy((unsigned)cand[i], x, y);
    }
    __device__ __forceinline__ int node(int i) const { return CACHED ? (int)s_node[i] : (int)cnode[i]; }
    __device__ __forceinline__ void set_node(int i, int v) const { if (CACHED) s_node[i] = (unsigned short)v; else cnode[i] = v; }
};

// Quadrant populations of a node list: four ints per node, or (PACK: every level holds at most 65535 candidates) four u16 in
// two words -- 16 bytes less shared memory per node and node list, which is what lets five CTAs share an SM at VGA size.
template <bool PACK> struct QCount {
    uint32_t *w;
    __device__ __forceinline__ void clear(int nodes, int tid) const {
        for (int i = tid; i < (PACK ? 2 : 4) * nodes; i += DIST_THREADS) w[i] = 0u;
    }
    __device__ __forceinline__ void add(int node, int q) const {
        if (PACK) atomicAdd(&w[2 * node + (q >> 1)], 1u << (16 * (q & 1)));
        else atomicAdd(&w[4 * node + q], 1u);
    }
    __device__ __forceinline__ int get(int node, int q) const {
        return PACK ? (int)((w[2 * node + (q >> 1)] >> (16 * (q & 1))) & 0xffffu) : (int)w[4 * node + q];
    }
    __device__ __forceinline__ int children(int node) const {   // quadrants that hold a candidate
        if (PACK) {
            const uint32_t a = w[2 * node], b = w[2 * node + 1];
            return ((a & 0xffffu) != 0u) + ((a >> 16) != 0u) + ((b & 0xffffu) != 0u) + ((b >> 16) != 0u);
        }
        return (w[4 * node] > 0u) + (w[4 * node + 1] > 0u) + (w[4 * node + 2] > 0u) + (w[4 * node + 3] > 0u);
    }
};

struct DistShared {
    unsigned long long *best;
    NodeBox *box0, *box1;
    int *cnt0, *cnt1, *ord, *byord, *ps, *kpos;
    uint32_t *ccnt0, *ccnt1;
    unsigned short *cpos;
};

template <bool CACHED, bool PACK>
__device__ void distribute_level(const LevelDev &L, int NC, int ncand, const CandStore<CACHED> &cs, const DistShared &sh, int *tmp,
                                 int *s_m, int *out_xy, int *out_resp, int *kp_count_out, int *err) {
    const int tid = threadIdx.x;
    const int N = L.budget;
    int flip = 0;
    NodeBox *box = sh.box0, *nbox = sh.box1;
    int *cnt = sh.cnt0, *ncnt = sh.cnt1;
    QCount<PACK> ccnt{sh.ccnt0}, nccnt{sh.ccnt1};    // quadrant populations of the current / the next node list
    unsigned short *cpos = sh.cpos;
    int *ord = sh.ord, *byord = sh.byord, *ps = sh.ps, *kpos = sh.kpos;

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
        cand_xy((unsigned)cs.cand[i], x, y);
        if (CACHED) cs.s_xy[i] = (uint32_t)x | ((uint32_t)y << 16);
        int nd = 0;
        if (n > 1) {
            const unsigned ix = (unsigned)(x / dx), iy = (unsigned)(y / dy);
            nd = min((int)(ix + iy * nx), n - 1);
        }
        cs.set_node(i, nd);
        if (n > 1) atomicAdd(&ncnt[nd], 1);
    }
    if (n == 1 && tid == 0) ncnt[0] = ncand;
    __syncthreads();
    // drop empty initial nodes, keeping their order
    if (n > 1) {
        for (int i = tid; i < n; i += DIST_THREADS) kpos[i] = ncnt[i] > 0 ? 1 : 0;
        __syncthreads();
        const int kept = block_exclusive_scan(kpos, n, tmp, flip);
        for (int i = tid; i < n; i += DIST_THREADS)
            if (ncnt[i] > 0) { box[kpos[i]] = nbox[i]; cnt[kpos[i]] = ncnt[i]; }
        ccnt.clear(kept, tid);
        __syncthreads();
        // (quadrant populations of the first round are counted on the way)
        for (int i = tid; i < ncand; i += DIST_THREADS) {
            const int nd = kpos[cs.node(i)];
            cs.set_node(i, nd);
            if (cnt[nd] > 1) {
                int x, y;
                cs.xy(i, x, y);
                ccnt.add(nd, quadrant(box[nd], x, y));
            }
        }
        n = kept;
        __syncthreads();
    } else {
        { NodeBox *tb = box; box = nbox; nbox = tb; int *tc = cnt; cnt = ncnt; ncnt = tc; }
        ccnt.clear(1, tid);
        __syncthreads();
        if (ncand > 1) {
            const NodeBox b0 = box[0];
            for (int i = tid; i < ncand; i += DIST_THREADS) {
                int x, y;
                cs.xy(i, x, y);
                ccnt.add(0, quadrant(b0, x, y));
            }
        }
        __syncthreads();
    }

    // ---- rounds ----------------------------------------------------------------------------------------
    bool partial = false;
    while (true) {
        // (A/B, the quadrant populations of every dividable node, were counted while the candidates moved: ccnt)
        // C: processing order of the dividable nodes
        int P;
        for (int j = tid; j < n; j += DIST_THREADS) ord[j] = cnt[j] > 1 ? 1 : 0;
        __syncthreads();
        P = block_exclusive_scan(ord, n, tmp, flip);       // whole round: list order
        if (partial) {
            // most populated first; equal counts: the node nearer the list front (the newer one) first
            if (n <= 4096 && ncand < (1 << 19)) {
                // rank = number of dividable nodes with a larger (count, -position) key; the keys of the P dividable
                // nodes are compacted first (ps as scratch), so the loop runs over P, not n, entries
                for (int j = tid; j < n; j += DIST_THREADS)
                    if (cnt[j] > 1) ps[ord[j]] = (cnt[j] << 12) | (4095 - j);
                __syncthreads();
                for (int j = tid; j < n; j += DIST_THREADS) {
                    if (cnt[j] > 1) {
                        const int kj = (cnt[j] << 12) | (4095 - j);
                        int r = 0;
                        for (int k = 0; k < P; ++k) r += ps[k] > kj ? 1 : 0;
                        ord[j] = r;                           // only thread j reads or writes ord[j] here
                    }
                }
            } else {                                            // keys would not fit 31 bits
                for (int j = tid; j < n; j += DIST_THREADS) {
                    const int cj = cnt[j];
                    int r = 0;
                    if (cj > 1)
                        for (int k = 0; k < n; ++k) {
                            const int ck = cnt[k];
                            r += (ck > 1 && (ck > cj || (ck == cj && k < j))) ? 1 : 0;
                        }
                    if (cj > 1) ord[j] = r;
                }
            }
            __syncthreads();
        }
        for (int j = tid; j < n; j += DIST_THREADS)
            if (cnt[j] > 1) byord[ord[j]] = j;
        __syncthreads();
        // D: children created before each processed node; in a partial round, where to stop
        for (int i = tid; i < P; i += DIST_THREADS) {
            ps[i] = ccnt.children(byord[i]);
        }
        if (tid == 0) { ps[P] = 0; *s_m = P; }
        __syncthreads();
        block_exclusive_scan(ps, P + 1, tmp, flip);   // ps[i] = children created before order index i; ps[P] = all
        if (partial) {
            // list size after processing order index i:  n + (ps[i+1] - (i+1))
            for (int i = tid; i < P; i += DIST_THREADS) {
                const bool reached = N <= n + ps[i + 1] - (i + 1);
                const bool before = i > 0 && N <= n + ps[i] - i;
                if (reached && !before) *s_m = i + 1;
            }
            __syncthreads();
        }
        const int m = *s_m;
        const int total_new = ps[m];
        const int n_new = total_new + (n - m);
        if (n_new > NC) { if (tid == 0) { atomicExch(err, SG_ERR_OVERFLOW); *kp_count_out = 0; } return; }
        // E: survivors keep their relative order behind the new children.  Whole round: every dividable node is processed,
        //    so the survivors before j are j minus the dividable nodes before j (ord holds exactly that count).
        if (partial) {
            for (int j = tid; j < n; j += DIST_THREADS) kpos[j] = (cnt[j] > 1 && ord[j] < m) ? 0 : 1;
            __syncthreads();
            block_exclusive_scan(kpos, n, tmp, flip);
        }
        // F: new node table (and cleared quadrant populations for it)
        nccnt.clear(n_new, tid);
        for (int j = tid; j < n; j += DIST_THREADS) {
            if (cnt[j] > 1 && ord[j] < m) {
                const NodeBox b = box[j];
                const int mx = b.bx + ((b.ex - b.bx + 1) >> 1), my = b.by + ((b.ey - b.by + 1) >> 1);
                int created = ps[ord[j]];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = ccnt.get(j, q);
                    if (c == 0) continue;
                    const int pos = total_new - 1 - created++;   // pushed to the front in creation order
                    NodeBox nb;
                    nb.bx = (q & 1) ? (short)mx : b.bx; nb.ex = (q & 1) ? b.ex : (short)mx;
                    nb.by = (q & 2) ? (short)my : b.by; nb.ey = (q & 2) ? b.ey : (short)my;
                    nbox[pos] = nb;
                    ncnt[pos] = c;
                    cpos[4 * j + q] = (unsigned short)pos;
                }
            } else {
                const int pos = total_new + (partial ? kpos[j] : j - ord[j]);
                nbox[pos] = box[j];
                ncnt[pos] = cnt[j];
                kpos[j] = pos;
            }
        }
        __syncthreads();
        // G: move the candidates; a candidate that lands in a dividable node is counted in that node's quadrant right
        //    away (A/B of the next round: no second pass over the candidates)
        const bool more = n_new < N && n_new != n;          // another round follows
        for (int i = tid; i < ncand; i += DIST_THREADS) {
            const int nd = cs.node(i);
            const bool moved = cnt[nd] > 1 && ord[nd] < m;
            if (!moved && !more) { cs.set_node(i, kpos[nd]); continue; }
            int x, y;
            cs.xy(i, x, y);
            const int nn = moved ? cpos[4 * nd + quadrant(box[nd], x, y)] : kpos[nd];
            cs.set_node(i, nn);
            if (more && ncnt[nn] > 1) nccnt.add(nn, quadrant(nbox[nn], x, y));
        }
        __syncthreads();
        // H: termination (uniform)
        const int n_old = n;
        n = n_new;
        { NodeBox *tb = box; box = nbox; nbox = tb; int *tc = cnt; cnt = ncnt; ncnt = tc; uint32_t *tw = ccnt.w; ccnt.w = nccnt.w; nccnt.w = tw; }
        if (N <= n || n == n_old) break;
        if (!partial) {
            // dividable nodes of the new list (all of them are children made in this round)
            const int pool = block_count(n, [&](int j) { return cnt[j] > 1; });
            if (N < n + 3 * pool) partial = true;
        }
    }

    // ---- strongest candidate of every node; earlier candidate wins ties --------------------------------
    unsigned long long *best = sh.best;
    for (int j = tid; j < n; j += DIST_THREADS) best[j] = 0ull;
    __syncthreads();
    for (int i = tid; i < ncand; i += DIST_THREADS) {
        const unsigned long long c = cs.cand[i];
        atomicMax(&best[cs.node(i)], (c & 0xffffffff00000000ull) | (0xffffffffu - (unsigned)c));
    }
    __syncthreads();
    for (int j = tid; j < n; j += DIST_THREADS) {
        const unsigned long long b = best[j];
        int x, y;
        cand_xy(0xffffffffu - (unsigned)b, x, y);
        out_xy[j] = (x + PATCH_RADIUS) | ((y + PATCH_RADIUS) << 16);
        out_resp[j] = (int)(b >> 32);
    }
    if (tid == 0) *kp_count_out = n;
}

template <bool PACK>
__global__ void __launch_bounds__(DIST_THREADS)
distribute_kernel(const __grid_constant__ GeomDev g, int node_cap_max, const unsigned long long *cand_all,
                  uint32_t *cand_node_all, const int *cand_count, int *kp_xy, int *kp_resp, int *kp_count,
                  int *err) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const int NC = node_cap_max;
    DistShared sh;
    sh.box0 = reinterpret_cast<NodeBox *>(dsm);                                    // [NC]
    sh.box1 = sh.box0 + NC;                                                        // [NC]
    sh.cnt0 = reinterpret_cast<int *>(sh.box1 + NC);                               // [NC]
    sh.cnt1 = sh.cnt0 + NC;                                                        // [NC]
    constexpr int QW = PACK ? 2 : 4;                                               // words of quadrant counts per node
    sh.ccnt0 = reinterpret_cast<uint32_t *>(sh.cnt1 + NC);                         // [QW NC] child counts (this round)
    sh.ccnt1 = sh.ccnt0 + QW * NC;                                                 // [QW NC] child counts (next round)
    sh.best = reinterpret_cast<unsigned long long *>(sh.ccnt0);                    // [NC] after the rounds, over the child counts
    sh.ord = reinterpret_cast<int *>(sh.ccnt1 + QW * NC);                          // [NC] processing index
    sh.byord = sh.ord + NC;                                                        // [NC]
    sh.ps = sh.byord + NC;                                                         // [NC + 1] scan over order
    sh.kpos = sh.ps + NC + 1;                                                      // [NC] position of kept nodes
    sh.cpos = reinterpret_cast<unsigned short *>(sh.kpos + NC + 1);                // [4 NC] child positions (NC < 65536)
    uint32_t *s_xy = reinterpret_cast<uint32_t *>(sh.cpos + 4 * NC + 4);           // [CAND_SMEM] x | y << 16
    unsigned short *s_node = reinterpret_cast<unsigned short *>(s_xy + CAND_SMEM); // [CAND_SMEM]
    __shared__ int tmp[16];
    __shared__ int s_m;

    // grid (frames, levels): CTAs are issued frame-fastest, so the long level-0 CTAs of all frames start first and the
    // short top-level ones fill the tail of the launch
    const int tid = threadIdx.x, l = blockIdx.y, f = blockIdx.x + g.frame0;
    const LevelDev &L = g.lv[l];
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
    if (ncand <= CAND_SMEM && NC < 65536) {
        const CandStore<true> cs{cand, cnode, s_xy, s_node};
        distribute_level<true, PACK>(L, NC, ncand, cs, sh, tmp, &s_m, out_xy, out_resp, kp_count + f * g.levels + l, err);
    } else {
        const CandStore<false> cs{cand, cnode, s_xy, s_node};
        distribute_level<false, PACK>(L, NC, ncand, cs, sh, tmp, &s_m, out_xy, out_resp, kp_count + f * g.levels + l, err);
    }
}

// {level | cell row << 8 | cell column << 20, first evaluated x, y, width | height << 16} of every FAST cell (host, sg_create)
void fast_cell_table(const GeomDev &g, std::vector<int4> &cells) {
    cells.clear();
    for (int l = 0; l < g.levels; ++l) {
        const LevelDev &L = g.lv[l];
        for (int ci = 0; ci < L.cells_y; ++ci)
            for (int cj = 0; cj < L.cells_x; ++cj) {
                const int ex0 = EVAL_ORIGIN + CELL * cj, ey0 = EVAL_ORIGIN + CELL * ci;
                const int cw = std::min(CELL, L.w - EVAL_ORIGIN - ex0), ch = std::min(CELL, L.h - EVAL_ORIGIN - ey0);
                cells.push_back(make_int4(l | (ci << 8) | (cj << 20), ex0, ey0, cw | (ch << 16)));
            }
    }
}

size_t distribute_smem_bytes(int node_cap_max, bool pack) {
    const size_t NC = node_cap_max;
    return NC * (2 * sizeof(NodeBox) + 2 * 4 + 2 * (pack ? 2 : 4) * 4 + 4 * 2 + 4 + 4 + 4 + 4) + 64 + (size_t)CAND_SMEM * 6;
}

// Upper bound of the NMS survivors of a level: one per 2 x 2 block of every cell.
static int max_candidates(const LevelDev &L) {
    int total = 0;
    for (int cj = 0; cj < L.cells_y; ++cj)
        for (int ci = 0; ci < L.cells_x; ++ci) {
            const int cw = std::min(CELL, L.w - 2 * EVAL_ORIGIN - CELL * ci), ch = std::min(CELL, L.h - 2 * EVAL_ORIGIN - CELL * cj);   // fast_cell_table
            total += ((cw + 1) / 2) * ((ch + 1) / 2);
        }
    return total;
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
        if (!ctx->fast_carveout_set) {          // 5 CTAs x 37 KB of static shared memory per SM need the largest carve-out
            SG_CUDA(ctx, cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            ctx->fast_carveout_set = true;
        }
        FastMaps maps;
        for (int l = 0; l < g.levels; ++l) maps.m[l] = ctx->lv[l].map_fast;
        fast_cells_kernel<<<dim3(total_cells, n_frames), FAST_THREADS, 0, ctx->stream>>>(
            g, maps, ctx->d_cell_table, ctx->d_cand, ctx->d_cand_count, ctx->d_err + ctx->err_slot);
        SG_LAUNCH_CHECK(ctx);
    }
    mark(ctx, EV_FAST1);
    if (nc_max >= 65536) return fail(ctx, SG_ERR_INVALID, "more than 65535 quadtree nodes per level are not supported");
    // quadrant counts as packed u16 when no level can hold more than 65535 candidates (VGA: 64 964 at level 0)
    bool pack = true;
    for (int l = 0; l < g.levels; ++l) pack = pack && max_candidates(g.lv[l]) <= 65535;
    const size_t smem = distribute_smem_bytes(nc_max, pack);
    auto kernel = pack ? distribute_kernel<true> : distribute_kernel<false>;
    if (smem > 48 * 1024)
        SG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!ctx->dist_carveout_set) {
        SG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ctx->dist_carveout_set = true;
    }
    kernel<<<dim3(n_frames, g.levels), DIST_THREADS, smem, ctx->stream>>>(
        g, nc_max, ctx->d_cand, ctx->d_cand_node, ctx->d_cand_count, ctx->d_kp_xy, ctx->d_kp_resp, ctx->d_kp_count, ctx->d_err + ctx->err_slot);
    SG_LAUNCH_CHECK(ctx);
    mark(ctx, EV_DIST1);
    ctx->detected = true;
    return SG_OK;
}

}  // namespace sg
