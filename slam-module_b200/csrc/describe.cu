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
#include "ctx.h"

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

__global__ void __launch_bounds__(DESC_WARPS * 32)
describe_kernel(const __grid_constant__ GeomDev g, const uint8_t *level0, int level0_pitch,
                unsigned long long level0_stride, const int *kp_xy, const int *kp_count,
                const int *trk_xy, const float *trk_pt, const int *trk_id, const int *trk_count,
                int track_level, DescOut o) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y + g.frame0;
    const int slot = blockIdx.x * DESC_WARPS + warp;
    const int n_trk = trk_count ? trk_count[f] : 0;

    // which keypoint is this slot: tracker points first, then level 0, 1, ... (orb_extractor.cpp:89-162)
    int l, ix, iy, tid_out = -1;
    float ox, oy;
    int total = n_trk;
    for (int k = 0; k < g.levels; ++k) total += kp_count[f * g.levels + k];
    if (slot == 0 && lane == 0) o.count[f] = total;
    if (slot >= total) return;
    if (slot < n_trk) {
        l = track_level;
        const int p = trk_xy[f * g.max_tracks + slot];
        ix = p & 0xffff; iy = p >> 16;
        ox = trk_pt[2 * (f * g.max_tracks + slot)];
        oy = trk_pt[2 * (f * g.max_tracks + slot) + 1];
        tid_out = trk_id[f * g.max_tracks + slot];
    } else {
        int t = slot - n_trk;
        for (l = 0; l < g.levels; ++l) {
            const int c = kp_count[f * g.levels + l];
            if (t < c) break;
            t -= c;
        }
        const int p = kp_xy[(size_t)f * g.det_cap + g.lv[l].kp_off + t];
        ix = p & 0xffff; iy = p >> 16;
        ox = __fmul_rn((float)ix, g.lv[l].scale);   // kp.pt * scale_at_level (orb_extractor.cpp:156)
        oy = __fmul_rn((float)iy, g.lv[l].scale);
    }
    const LevelDev &L = g.lv[l];
    const uint8_t *img = l == 0 ? level0 + (size_t)f * level0_stride : L.pyr + (size_t)f * L.frame_stride;
    const int pitch = l == 0 ? level0_pitch : L.pitch;
    const uint8_t *blur = L.blur + (size_t)f * L.frame_stride;

    // ---- intensity-centroid moments over the radius-15 disc: lane = column u ---------------------------
    int m10 = 0, m01 = 0;
    {
        const int u = lane - HALF_PATCH;
        const uint8_t *c = img + (size_t)iy * pitch + ix;
        if (lane < 31) {
            const int au = abs(u);
#pragma unroll 1
            for (int v = -HALF_PATCH; v <= HALF_PATCH; ++v) {
                if (au <= c_umax[abs(v)]) {
                    const int val = __ldg(c + v * pitch + u);
                    m10 += u * val;
                    m01 += v * val;
                }
            }
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rBRIEF: lane = descriptor byte, 8 point pairs each ---------------------------------------------
    const float rad = (float)((double)angle * 3.14159265358979323846 / 180.0);
    const float cs = util_cos(rad), sn = util_sin(rad);
    const uint8_t *cb = blur + (size_t)iy * L.pitch + ix;
    const int4 *pat = reinterpret_cast<const int4 *>(d_pattern) + 2 * lane;
    unsigned byte = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int4 w = __ldg(pat + h);
        const int ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x0 = (float)(int8_t)(ws[k] & 0xff), y0 = (float)(int8_t)((ws[k] >> 8) & 0xff);
            const float x1 = (float)(int8_t)((ws[k] >> 16) & 0xff), y1 = (float)(int8_t)((ws[k] >> 24) & 0xff);
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(y0, cs)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(y0, sn)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sn), __fmul_rn(y1, cs)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, cs), __fmul_rn(y1, sn)));
            const int v0 = __ldg(cb + r0 * L.pitch + c0), v1 = __ldg(cb + r1 * L.pitch + c1);
            byte |= (v0 < v1 ? 1u : 0u) << (4 * h + k);
        }
    }
    unsigned word = byte << (8 * (lane & 3));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);

    const size_t oi = (size_t)f * g.out_cap + slot;
    if ((lane & 3) == 0) o.desc[8 * oi + (lane >> 2)] = word;
    if (lane == 0) {
        o.x[oi] = ox; o.y[oi] = oy; o.angle[oi] = angle; o.octave[oi] = l;
        o.track_id[oi] = tid_out; o.lvl_x[oi] = ix; o.lvl_y[oi] = iy;
    }
}

int launch_describe(sg_ctx *ctx, int n_frames) {
    GeomDev g = ctx->geom;
    g.frame0 = ctx->frame0;
    DescOut o{ctx->d_x, ctx->d_y, ctx->d_angle, ctx->d_octave, ctx->d_track_id, ctx->d_lvl_x, ctx->d_lvl_y,
              ctx->d_desc, ctx->d_count};
    const bool trk = ctx->have_tracks;
    dim3 grid((g.out_cap + DESC_WARPS - 1) / DESC_WARPS, n_frames);
    describe_kernel<<<grid, DESC_WARPS * 32, 0, ctx->stream>>>(
        g, ctx->level0, ctx->level0_pitch, ctx->level0_stride, ctx->d_kp_xy, ctx->d_kp_count,
        trk ? ctx->d_trk_xy : nullptr, trk ? ctx->d_trk_pt : nullptr, trk ? ctx->d_trk_id : nullptr,
        trk ? ctx->d_trk_count : nullptr, ctx->p.track_level, o);
    SG_LAUNCH_CHECK(ctx);
    mark(ctx, EV_DESC1);
    return SG_OK;
}

}  // namespace sg
