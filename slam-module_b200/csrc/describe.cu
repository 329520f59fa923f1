// Orientation (intensity centroid) and rotated-BRIEF descriptors, one warp per keypoint.
//
// Replaces ic_angle (orb_extractor.cpp:245-275), compute_orb_descriptor (orb_extractor.cpp:284-352,
// the non-SSE branch :326-331 that the reference build uses), util::cos / util::sin
// (openvslam/trigonometric.h:17-46), cv::fastAtan2 (OpenCV core, scalar path) and the output
// assembly of detectAndExtract (orb_extractor.cpp:89-124 tracker points, :153-162 detected points).
//
// Bit-exactness: moments are int32 sums (order-free); every fp32 step uses explicit
// round-to-nearest mul / add / sub / div intrinsics in the reference's evaluation order (no FMA
// contraction), cvRound is round-half-even (__float2int_rn), the degree -> radian conversion is
// done in double like orb_extractor.cpp:286.
#include <algorithm>
#include "ctx.h"
#include "tma.cuh"

namespace sg {

constexpr int DESC_WARPS = 8;

// 256 point pairs (x0, y0, x1, y1) as int8 (openvslam/orb_point_pairs.h:47-304 holds the same integers as
// floats).  Global memory, not __constant__: every lane reads its own 32 bytes (coalesced), which the
// constant cache would serialise.
__device__ __align__(16) int8_t d_pattern[1024] = {
#include "orb_pattern.inc"
};
// u_max_ of orb_extractor.cpp:174-186 for a half patch size of 15
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ float poly_cos(float v) {
    const float c1 = 0.99940307f, c2 = -0.49558072f, c3 = 0.03679168f;
    const float v2 = __fmul_rn(v, v);
    return __fadd_rn(c1, __fmul_rn(v2, __fadd_rn(c2, __fmul_rn(c3, v2))));
}

__device__ __forceinline__ float util_cos(float v) {
    const float PI = 3.14159265358979f, PI_2 = PI / 2.0f, TWO_PI = 2.0f * PI, INV_TWO_PI = 1.0f / TWO_PI;
    const float THREE_PI_2 = 3.0f * PI_2;
    const float q = __fmul_rn(v, INV_TWO_PI);
    int fl = (int)q;            // cvFloor
    fl -= (fl > q) ? 1 : 0;
    v = __fsub_rn(v, __fmul_rn((float)fl, TWO_PI));
    v = (0.0f < v) ? v : -v;
    if (v < PI_2) return poly_cos(v);
    if (v < PI) return -poly_cos(__fsub_rn(PI, v));
    if (v < THREE_PI_2) return -poly_cos(__fsub_rn(v, PI));
    return poly_cos(__fsub_rn(TWO_PI, v));
}
__device__ __forceinline__ float util_sin(float v) {
    const float PI_2 = 3.14159265358979f / 2.0f;
    return util_cos(__fsub_rn(PI_2, v));
}

struct DescOut {
    float *x, *y, *angle;
    int *octave, *track_id, *lvl_x, *lvl_y;
    uint32_t *desc;
    int *count;
};

// ---- kernel geometry ------------------------------------------------------------------------------------
// A warp owns a GROUP of 32 consecutive output slots of one frame and runs three phases over it:
//   A  moments, warp-cooperative per keypoint: 3 patch rows per step, lane = (row, aligned word), the row
//      sums  sum(u*I)  and  sum(I)  are one IDP.4A each against per-(alignment, |v|) weight words in shared
//      memory (the disc mask u_max_ is folded into the weights);
//   B  lane j finishes keypoint j: fastAtan2, degree -> radian (double), util::cos / util::sin, and the
//      coalesced stores of x, y, angle, octave, track id, level coordinates;
//   C  descriptor, warp-cooperative per keypoint: the 37x37 window of the blurred plane is staged in shared
//      memory with aligned word loads, lane = descriptor byte samples its 8 point pairs from there; the 32
//      pattern floats of a lane live in registers for the whole (persistent) kernel; cvRound is the
//      1.5*2^23 magic-number add (exact round-half-even for |v| < 2^22) instead of the quarter-rate F2I.
constexpr int MOM_WORDS = 9;            // aligned words covering the 31 columns of the moment disc
constexpr int BLUR_R = 18;              // max |rotated pattern coordinate| (pattern radius 18.38, |cos|,|sin| <= 1.001)
// TMA boxes 64 x 37 / 48 x 31 with the x origin aligned down to 16 bytes: a box origin at the exact window column
// (48 x 37 / 32 x 31 boxes, 28% fewer bytes) raises "illegal instruction" on sm_100a / driver 580 -- measured twice.
constexpr int XALIGN = 16, BLUR_PITCH = 64, MOM_PITCH = 48;
constexpr int BLUR_ROWS = 2 * BLUR_R + 1, MOM_ROWS = 2 * HALF_PATCH + 1;
constexpr int PATCH_BYTES = 2560;                            // per buffer: >= 37 * 64 and >= 31 * 48, multiple of 512 (swizzle period)
constexpr float ROUND_MAGIC = 12582912.f;        // 1.5 * 2^23
constexpr int ROUND_MAGIC_BITS = 0x4B400000;

struct KpInfo { int ix, iy, l, tid; float ox, oy; };

// unsigned pixel bytes x signed weight bytes (IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// packed fp32 pairs of sm_100 (FMUL2 / FADD2): every element is an IEEE round-to-nearest operation of its own
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_f32(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi_f32(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

struct DescMaps { CUtensorMap mom[SG_MAX_LEVELS], blur[SG_MAX_LEVELS]; };   // 48 x 31 boxes over the pyramid planes, 64 x 37 over the blurred ones

__global__ void __launch_bounds__(DESC_WARPS * 32)
describe_kernel(const __grid_constant__ GeomDev g, const __grid_constant__ DescMaps maps, const int *kp_xy, const int *kp_count,
                const int *trk_xy, const float *trk_pt, const int *trk_id, const int *trk_count,
                int track_level, int n_frames, int split, DescOut o) {
    // moment weights per (alignment of the window inside its word, |v|, word): the signed byte u and the byte |v| inside the
    // disc, 0 outside, so that m10 and the two halves of m01 are two IDP.4A per pixel word with no other arithmetic
    __shared__ uint32_t s_w[2][4][16][MOM_WORDS];      // [0]: u, [1]: |v|
    __shared__ __align__(512) uint8_t s_patch[DESC_WARPS][2 * PATCH_BYTES];   // two TMA destinations per warp, 512-byte aligned
    __shared__ __align__(8) uint64_t s_bar[DESC_WARPS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- per-CTA tables --------------------------------------------------------------------------------
    for (int i = threadIdx.x; i < 4 * 16 * MOM_WORDS; i += DESC_WARPS * 32) {
        const int w = i % MOM_WORDS, av = (i / MOM_WORDS) % 16, off = i / (MOM_WORDS * 16);
        uint32_t wu = 0, wv = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = 4 * w + j - HALF_PATCH - off;
            if (abs(u) <= c_umax[av]) { wu |= (uint32_t)(uint8_t)(int8_t)u << (8 * j); wv |= (uint32_t)av << (8 * j); }
        }
        (&s_w[0][0][0][0])[i] = wu;
        (&s_w[1][0][0][0])[i] = wv;
    }
    // Both points of a pair share one packed register: {x0, x1} and {y0, y1} (FMUL2 / FADD2 work on two floats at once).
    unsigned long long pxx[8], pyy[8];
    {
        const int4 *pat = reinterpret_cast<const int4 *>(d_pattern) + 2 * lane;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int4 w = __ldg(pat + h);
            const int ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                pxx[4 * h + k] = pack_f32x2((float)(int8_t)(ws[k] & 0xff), (float)(int8_t)((ws[k] >> 16) & 0xff));
                pyy[4 * h + k] = pack_f32x2((float)(int8_t)((ws[k] >> 8) & 0xff), (float)(int8_t)((ws[k] >> 24) & 0xff));
            }
        }
    }
    if (lane == 0) { mbar_init(&s_bar[warp][0], 1); mbar_init(&s_bar[warp][1], 1); }
    __syncthreads();
    unsigned bar_phase = 0;    // bit b: parity the next wait on buffer b expects

    const int groups_per_frame = (g.out_cap + 31) >> 5;
    const int n_groups = groups_per_frame * n_frames;
    const int mrow = lane / MOM_WORDS, mword = lane - mrow * MOM_WORDS;        // lanes 0..26: 3 rows x 9 words
    uint8_t *patch = s_patch[warp];
    uint64_t *bars = s_bar[warp];

    // small batches (single-frame calls): a group is cut into `split` parts of 32 / split keypoints, one warp each, so
    // that the grid still fills the chip; every warp of a group does the (cheap) slot lookup for all 32 lanes
    const int per_part = 32 / split;
    for (int vg = blockIdx.x * DESC_WARPS + warp; vg < n_groups * split; vg += gridDim.x * DESC_WARPS) {
        const int grp = vg / split, part = vg - grp * split;
        const int f = g.frame0 + grp / groups_per_frame;
        const int slot0 = (grp % groups_per_frame) << 5;
        const int n_trk = trk_count ? trk_count[f] : 0;
        // which keypoints: tracker points first, then level 0, 1, ... (orb_extractor.cpp:89-162)
        int cnt = lane < g.levels ? kp_count[f * g.levels + lane] : 0;
        int incl = cnt;                                    // inclusive scan over the levels
#pragma unroll
        for (int d = 1; d < SG_MAX_LEVELS; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int total = n_trk + __shfl_sync(0xffffffffu, incl, SG_MAX_LEVELS - 1);
        if (slot0 == 0 && lane == 0 && part == 0) o.count[f] = total;
        if (slot0 >= total) continue;
        const int n_here = min(32, total - slot0);
        const int jlo = part * per_part, jhi = min(n_here, jlo + per_part);
        if (jlo >= jhi) continue;

        // ---- lane j looks up keypoint slot0 + j -------------------------------------------------------------
        KpInfo k{0, 0, 0, -1, 0.f, 0.f};
        const int slot = slot0 + lane;
        {
            const int t = slot - n_trk;
            // level of a detected keypoint: number of levels whose inclusive count is <= t
            int l = 0;
#pragma unroll
            for (int q = 0; q < SG_MAX_LEVELS; ++q) {
                const int iq = __shfl_sync(0xffffffffu, incl, q);
                l += (q < g.levels && iq <= t) ? 1 : 0;
            }
            if (slot < total) {
                if (slot < n_trk) {
                    k.l = track_level;
                    const int p = trk_xy[f * g.max_tracks + slot];
                    k.ix = p & 0xffff; k.iy = p >> 16;
                    k.ox = trk_pt[2 * (f * g.max_tracks + slot)];
                    k.oy = trk_pt[2 * (f * g.max_tracks + slot) + 1];
                    k.tid = trk_id[f * g.max_tracks + slot];
                } else {
                    k.l = l;
                }
            }
        }
        // exclusive count of the lane's level (second pass: shuffles must be warp-uniform)
        {
            const int lsel = (slot < total && slot >= n_trk) ? k.l : 0;
            const int excl = __shfl_sync(0xffffffffu, incl - cnt, lsel);
            if (slot < total && slot >= n_trk) {
                const int t = slot - n_trk - excl;
                const int p = kp_xy[(size_t)f * g.det_cap + g.lv[k.l].kp_off + t];
                k.ix = p & 0xffff; k.iy = p >> 16;
                k.ox = __fmul_rn((float)k.ix, g.lv[k.l].scale);   // kp.pt * scale_at_level (orb_extractor.cpp:156)
                k.oy = __fmul_rn((float)k.iy, g.lv[k.l].scale);
            }
        }

        // ---- phase A: intensity-centroid moments (un-blurred level) ----------------------------------------
        // double buffered: lane 0 issues the TMA box load (48 x 31, x origin aligned down to 16 bytes) of keypoint
        // j + 1 while the warp sums keypoint j from shared memory
        int my_m10 = 0, my_m01 = 0;
        {
            auto issue = [&](int j) {       // lane j owns keypoint j: it issues the load itself (no shuffles of the box origin)
                if (lane == j) {
                    uint64_t *bar = bars + (j & 1);
                    mbar_expect_tx(bar, MOM_ROWS * MOM_PITCH);
                    tma_load_3d(patch + (j & 1) * PATCH_BYTES, &maps.mom[k.l], (k.ix - HALF_PATCH) & ~(XALIGN - 1), k.iy - HALF_PATCH, f, bar);
                }
            };
            __syncwarp();   // every lane is done with both buffers (previous group)
            issue(jlo);
            for (int j = jlo; j < jhi; ++j) {
                const int b = j & 1;
                if (j + 1 < jhi) issue(j + 1);        // buffer b ^ 1 was last read for keypoint j - 1 (syncwarp below)
                const int ix = __shfl_sync(0xffffffffu, k.ix, j);
                const int off16 = (ix - HALF_PATCH) & (XALIGN - 1), off = off16 & 3;
                mbar_wait(bars + b, (bar_phase >> b) & 1u);
                bar_phase ^= 1u << b;
                int m10 = 0, m01 = 0;
                if (lane < 3 * MOM_WORDS) {
                    // lane = (row mod 3, word): window rows mrow, mrow + 3, ...; v = row - 15 is negative for i < 5 (|v| = 15 -
                    // mrow - 3 i) and mrow + 3 (i - 5) from there on, so every load address is a lane base plus an immediate
                    const uint32_t pix_a = smem_u32(patch + b * PATCH_BYTES + mrow * MOM_PITCH + (off16 & ~3) + 4 * mword);
                    const uint32_t up_a = smem_u32(&s_w[0][off][HALF_PATCH - mrow][mword]), dn_a = smem_u32(&s_w[0][off][mrow][mword]);
                    constexpr uint32_t WV = sizeof(s_w[0]);      // from a u word to the |v| word of the same (alignment, row, word)
                    int m01n = 0;
#pragma unroll
                    for (int i = 0; i < 11; ++i) {
                        if (3 * i + mrow < MOM_ROWS) {       // false only for i = 10, mrow > 0
                            const uint32_t pix = lds32(pix_a + 3 * i * MOM_PITCH);
                            const uint32_t wa = i < 5 ? up_a - 3 * i * MOM_WORDS * 4 : dn_a + 3 * (i - 5) * MOM_WORDS * 4;
                            m10 = dp4a_us(pix, (int)lds32(wa), m10);
                            if (i < 5) m01n = (int)__dp4a(pix, lds32(wa + WV), (unsigned)m01n);
                            else m01 = (int)__dp4a(pix, lds32(wa + WV), (unsigned)m01);
                        }
                    }
                    m01 -= m01n;
                }
                m10 = __reduce_add_sync(0xffffffffu, m10);     // (also orders this keypoint's reads before the next issue)
                m01 = __reduce_add_sync(0xffffffffu, m01);
                if (lane == j) { my_m10 = m10; my_m01 = m01; }
            }
        }

        // ---- phase B: lane j finishes keypoint j ---------------------------------------------------------------
        const float angle = fast_atan2_deg((float)my_m01, (float)my_m10);
        const float rad = (float)((double)angle * 3.14159265358979323846 / 180.0);
        const float cs = util_cos(rad), sn = util_sin(rad);
        const size_t oi0 = (size_t)f * g.out_cap + slot0;
        if (lane >= jlo && lane < jhi) {
            const size_t oi = oi0 + lane;
            o.x[oi] = k.ox; o.y[oi] = k.oy; o.angle[oi] = angle; o.octave[oi] = k.l;
            o.track_id[oi] = k.tid; o.lvl_x[oi] = k.ix; o.lvl_y[oi] = k.iy;
        }

        // ---- phase C: rBRIEF on the blurred level -----------------------------------------------------------
        // same double buffering with 64 x 37 boxes of the blurred plane
        auto issue_blur = [&](int j) {
            if (lane == j) {
                uint64_t *bar = bars + (j & 1);
                mbar_expect_tx(bar, BLUR_ROWS * BLUR_PITCH);
                tma_load_3d(patch + (j & 1) * PATCH_BYTES, &maps.blur[k.l], (k.ix - BLUR_R) & ~(XALIGN - 1), k.iy - BLUR_R, f, bar);
            }
        };
        __syncwarp();       // phase A's reads are complete
        issue_blur(jlo);
        for (int j = jlo; j < jhi; ++j) {
            const int b = j & 1;
            __syncwarp();   // keypoint j - 1 has been sampled by every lane: its buffer may be refilled
            if (j + 1 < jhi) issue_blur(j + 1);
            const int ix = __shfl_sync(0xffffffffu, k.ix, j);
            const float c_ = __shfl_sync(0xffffffffu, cs, j), s_ = __shfl_sync(0xffffffffu, sn, j);
            const int off = (ix - BLUR_R) & (XALIGN - 1);
            mbar_wait(bars + b, (bar_phase >> b) & 1u);
            bar_phase ^= 1u << b;
            // The blurred box is stored with TMA's 64-byte swizzle: byte i = row * 64 + column of the box lives at
            // i ^ ((i >> 3) & 0x30) (16-byte chunk index XOR bits 7-8 of the offset; the buffer is 512-byte aligned, so the
            // shared address itself can be swizzled).  Un-swizzled, bank = 16 * (row & 1) + column / 4 and the 32 samples of a
            // warp instruction pile up on the few banks under the patch centre (5.5 wavefronts per LDS measured); swizzled,
            // eight consecutive rows of one column chunk use eight different bank groups.
            // sample offset = (r + 18) * 64 + (c + 18 + off); r, c arrive as MAGIC_BITS + integer
            // (unsigned arithmetic: the magic offsets cancel modulo 2^32)
            const unsigned bias = smem_u32(patch + b * PATCH_BYTES) + (unsigned)(BLUR_R * BLUR_PITCH + BLUR_R + off)
                                  - (unsigned)ROUND_MAGIC_BITS * (unsigned)(BLUR_PITCH + 1);
            auto sample = [&](unsigned r_bits, unsigned c_bits) {
                const unsigned i = r_bits * BLUR_PITCH + c_bits + bias;
                return (int)lds_u8(i ^ ((i >> 3) & 0x30u));
            };
            // rotation (orb_extractor.cpp:31-45): row = x * sin + y * cos, column = x * cos - y * sin, every product and sum
            // rounded to float on its own.  The products of both points run as packed FMUL2; the sums stay scalar FADDs
            // (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 -- one rounding too few -- even with --fmad=false);
            // x * cos - y * sin is computed as x * cos + y * (-sin), the same value bit for bit.
            const unsigned long long cc = pack_f32x2(c_, c_), ss = pack_f32x2(s_, s_), ns = pack_f32x2(-s_, -s_);
            const unsigned long long magic2 = pack_f32x2(ROUND_MAGIC, ROUND_MAGIC);
            unsigned bits = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned long long xs = mul_f32x2(pxx[q], ss), yc = mul_f32x2(pyy[q], cc);
                const unsigned long long xc = mul_f32x2(pxx[q], cc), yn = mul_f32x2(pyy[q], ns);
                const unsigned long long rr = add_f32x2(pack_f32x2(__fadd_rn(lo_f32(xs), lo_f32(yc)), __fadd_rn(hi_f32(xs), hi_f32(yc))), magic2);
                const unsigned long long cl = add_f32x2(pack_f32x2(__fadd_rn(lo_f32(xc), lo_f32(yn)), __fadd_rn(hi_f32(xc), hi_f32(yn))), magic2);
                const int v0 = sample((unsigned)rr, (unsigned)cl), v1 = sample((unsigned)(rr >> 32), (unsigned)(cl >> 32));
                bits |= (v0 < v1 ? 1u : 0u) << q;
            }
            unsigned word = bits << (8 * (lane & 3));
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            // 8 words per descriptor: lanes 0, 4, 8, ... hold words 0..7
            const unsigned wsel = __shfl_sync(0xffffffffu, word, (lane & 7) << 2);
            if (lane < 8) o.desc[8 * (oi0 + j) + lane] = wsel;
        }
    }
}

// TMA box extents of the describe kernel's windows (host, for the tensor maps)
void describe_box_dims(int *mom_w, int *mom_h, int *blur_w, int *blur_h) {
    *mom_w = MOM_PITCH; *mom_h = MOM_ROWS; *blur_w = BLUR_PITCH; *blur_h = BLUR_ROWS;
}

int launch_describe(sg_ctx *ctx, int n_frames) {
    GeomDev g = ctx->geom;
    g.frame0 = ctx->frame0;
    DescOut o{ctx->d_x, ctx->d_y, ctx->d_angle, ctx->d_octave, ctx->d_track_id, ctx->d_lvl_x, ctx->d_lvl_y,
              ctx->d_desc, ctx->d_count};
    const bool trk = ctx->have_tracks;
    const int groups = ((g.out_cap + 31) / 32) * n_frames;
    if (!ctx->describe_ctas_per_sm) {   // resident CTAs per SM of this (persistent) kernel, asked once per context
        int per_sm = 0;
        SG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, describe_kernel, DESC_WARPS * 32, 0));
        ctx->describe_ctas_per_sm = std::max(per_sm, 1);
    }
    const int resident_warps = ctx->sm_count * ctx->describe_ctas_per_sm * DESC_WARPS;
    int split = 1;
    while (split < 8 && groups * split * 2 <= resident_warps) split *= 2;
    const int blocks = std::max(1, std::min((groups * split + DESC_WARPS - 1) / DESC_WARPS, ctx->sm_count * ctx->describe_ctas_per_sm));
    DescMaps maps;
    for (int l = 0; l < g.levels; ++l) { maps.mom[l] = ctx->lv[l].map_mom; maps.blur[l] = ctx->lv[l].map_blur; }
    describe_kernel<<<blocks, DESC_WARPS * 32, 0, ctx->stream>>>(
        g, maps, ctx->d_kp_xy, ctx->d_kp_count,
        trk ? ctx->d_trk_xy : nullptr, trk ? ctx->d_trk_pt : nullptr, trk ? ctx->d_trk_id : nullptr,
        trk ? ctx->d_trk_count : nullptr, ctx->p.track_level, n_frames, split, o);
    SG_LAUNCH_CHECK(ctx);
    mark(ctx, EV_DESC1);
    return SG_OK;
}

}  // namespace sg
