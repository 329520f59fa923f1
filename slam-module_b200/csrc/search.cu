// "Next" rows of the hot path: the candidate-list matchers and the descriptor medoid.
//
//   FeatureSearch                 feature_search.cpp:22-48   Y-sorted index + radius query
//   searchByProjection            keyframe_matcher.cpp:356-386  best / second best over the query result, level-aware
//                                                             ratio rule, accepted keypoints are consumed
//   replaceDuplication            keyframe_matcher.cpp:482-499  best only, threshold 50
//   findMatchesTranformedMps      keyframe_matcher.cpp:604-627  best only, threshold 100, octave window [pred-1, pred]
//   MapPoint::updateDescriptor    map_point.cpp:75-116       Hamming medoid (median-of-distances arg-min)
//
// The geometry in front of the loops (reprojection, viewing distance, scale prediction) is the caller's: the
// kernels take the projected point, the search radius and the query descriptor.
//
// GPU formulation (exact):
//   search_topk_kernel    one warp per query: two binary searches bound the Y range, lanes test the circle
//                         (fp32 mul / add without contraction, like the x86 reference), compute the Hamming
//                         distance of the eligible candidates and keep the 4 smallest (distance, sorted position)
//                         keys of the warp -- ties resolve to the earlier position, i.e. the reference's scan order.
//   search_resolve_kernel mode 0: one thread per query (no interaction between queries).
//                         mode 1: one warp walks the queries in order; best / second are the first two unconsumed
//                         keys; when a truncated list cannot decide, the warp rescans the query's range exactly.
//   medoid_kernel         one CTA per map point: descriptors in shared memory, one warp per row of the distance
//                         matrix, the row median by bisection on the value range 0..256 with warp-wide counts.
#include <algorithm>
#include <vector>
#include "ctx.h"

extern "C" int sg_feature_index(const float *h_x, const float *h_y, int n, int32_t *h_order);

namespace sg {

constexpr int SK = 4;                       // keys kept per query
constexpr unsigned NOKEY = 0xffffffffu;
constexpr int SEARCH_WARPS = 8;
constexpr int MED_THREADS = 128, MED_MAX = 1024;   // descriptors per map point the medoid kernel stages

struct SearchArgs {
    const float *sx, *sy;        // keypoint coordinates in Y-sorted order
    const int *sidx;             // keypoint index of every sorted position
    const int *soct;             // octave of every sorted position
    const uint32_t *sdesc;       // descriptors in sorted order
    const unsigned char *elig0;  // per sorted position: 1 = not eligible from the start (mode 1: already matched)
    int nK;
    const float *qx, *qy, *qr;
    const uint32_t *qdesc;
    const int *qlevel;           // predicted scale level per query or nullptr (octave window off)
    int nQ, mode;
    unsigned thr;
};

__device__ __forceinline__ unsigned hamming8(const uint32_t (&a)[8], const uint32_t *b) {
    const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(b)), b1 = __ldg(reinterpret_cast<const uint4 *>(b) + 1);
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w)
           + __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// first position with !(sy[p] < v)
__device__ __forceinline__ int lower_bound_y(const float *sy, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sy + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first position with sy[p] > v   (the scan stops at the first element with !(y <= v))
__device__ __forceinline__ int upper_bound_y(const float *sy, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sy + mid) <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool in_circle(float qx, float qy, float r, float x, float y) {
    const float dx = __fsub_rn(qx, x), dy = __fsub_rn(qy, y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(r, r);   // feature_search.cpp:42-44
}

__device__ __forceinline__ bool level_ok(const SearchArgs &a, int q, int oct) {
    if (!a.qlevel) return true;
    const int pl = a.qlevel[q];
    return !(oct < pl - 1 || oct > pl);                                          // keyframe_matcher.cpp:611
}

__device__ __forceinline__ void insert4(unsigned (&t)[SK], unsigned key) {
#pragma unroll
    for (int i = 0; i < SK; ++i) {
        const unsigned lo = min(t[i], key);
        key = max(t[i], key);
        t[i] = lo;
    }
}

__global__ void __launch_bounds__(SEARCH_WARPS * 32)
search_topk_kernel(const SearchArgs a, uint32_t *keys, int2 *range, uint32_t *nseen_out) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * SEARCH_WARPS + (threadIdx.x >> 5);
    if (q >= a.nQ) return;
    const float qx = a.qx[q], qy = a.qy[q], r = a.qr[q];
    const int lo = lower_bound_y(a.sy, a.nK, __fsub_rn(qy, r)), hi = upper_bound_y(a.sy, a.nK, __fadd_rn(qy, r));
    uint32_t d[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) d[w] = __ldg(a.qdesc + 8 * (size_t)q + w);
    unsigned t[SK] = {NOKEY, NOKEY, NOKEY, NOKEY};
    unsigned nseen = 0;
    for (int p = lo + lane; p < hi; p += 32) {
        if (a.mode == 1 && a.elig0[p]) continue;
        if (!in_circle(qx, qy, r, __ldg(a.sx + p), __ldg(a.sy + p))) continue;
        if (!level_ok(a, q, __ldg(a.soct + p))) continue;
        ++nseen;
        insert4(t, (hamming8(d, a.sdesc + 8 * (size_t)p) << 16) | (unsigned)p);
    }
    nseen = __reduce_add_sync(0xffffffffu, nseen);
    // the 4 smallest keys of the warp: pop the minimum head four times
    unsigned out[SK];
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        const unsigned m = __reduce_min_sync(0xffffffffu, t[0]);
        out[k] = m;
        if (m != NOKEY && t[0] == m) { t[0] = t[1]; t[1] = t[2]; t[2] = t[3]; t[3] = NOKEY; }   // keys are unique (position)
    }
    if (lane == 0) {
        reinterpret_cast<uint4 *>(keys)[q] = make_uint4(out[0], out[1], out[2], out[3]);
        range[q] = make_int2(lo, hi);
        nseen_out[q] = nseen;
    }
}

__global__ void search_resolve_independent_kernel(const SearchArgs a, const uint32_t *keys, int *out_idx, unsigned *out_dist,
                                                  unsigned *n_matched) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.nQ) return;
    const unsigned k0 = keys[4 * (size_t)q];
    int idx = -1;
    unsigned dist = 256;
    if (k0 != NOKEY && (k0 >> 16) <= a.thr) { idx = a.sidx[k0 & 0xffffu]; dist = k0 >> 16; }
    out_idx[q] = idx;
    out_dist[q] = dist;
    if (idx >= 0) atomicAdd(n_matched, 1u);
}

// searchByProjection semantics: queries in order, accepted keypoints are consumed (keyframe_matcher.cpp:356-386).
// One warp, 32 queries at a time, optimistically (same scheme as match_resolve_kernel): every lane evaluates its
// query against the current `taken` bits and claims the keypoint it would take; the prefix of queries none of
// whose keys was claimed by an earlier query of the batch is committed at once -- its decisions are exactly the
// sequential ones -- and the first unsafe query is re-evaluated in the next round, or rescanned exactly by the
// whole warp when its truncated list cannot decide.
constexpr int SEARCH_CLAIMS = 2048;
__global__ void __launch_bounds__(32)
search_resolve_sequential_kernel(const SearchArgs a, const uint32_t *keys, const int2 *range, const uint32_t *nseen,
                                 unsigned char *taken_pos, int *out_idx, unsigned *out_dist, unsigned *n_matched,
                                 unsigned long long *rescans) {
    __shared__ uint32_t taken[2048];            // one bit per sorted position (nK <= 65535)
    __shared__ uint32_t claim[SEARCH_CLAIMS];   // slot = position mod SEARCH_CLAIMS: round << 8 | 31 - lane of the earliest claimant
    const int lane = threadIdx.x;
    // bitmap of the initially taken positions: one ballot per 32 positions
    for (int w0 = 0; w0 * 32 < a.nK; ++w0) {
        const int pos = 32 * w0 + lane;
        const unsigned bits = __ballot_sync(0xffffffffu, pos < a.nK && taken_pos[pos] != 0);
        if (lane == 0) taken[w0] = bits;
    }
    for (int i = lane; i < SEARCH_CLAIMS; i += 32) claim[i] = 0;
    __syncwarp();
    unsigned count = 0, n_rescan = 0, round_tag = 0;
    for (int base = 0; base < a.nQ; base += 32) {
        const int q = base + lane;
        uint4 kq = make_uint4(NOKEY, NOKEY, NOKEY, NOKEY);
        unsigned ns = 0;
        if (q < a.nQ) { kq = reinterpret_cast<const uint4 *>(keys)[q]; ns = nseen[q]; out_idx[q] = -1; out_dist[q] = 256; }
        const unsigned kk[SK] = {kq.x, kq.y, kq.z, kq.w};
        const bool complete = ns <= SK;
        unsigned pending = __ballot_sync(0xffffffffu, ns > 0);
        while (pending) {
            ++round_tag;
            const bool mine = (pending >> lane) & 1u;
            int decision = 0;   // 0 reject, 1 accept u0, 2 rescan
            unsigned u0 = NOKEY, u1 = NOKEY, last = NOKEY;
            if (mine) {
#pragma unroll
                for (int e = 0; e < SK; ++e) {
                    if (kk[e] == NOKEY) continue;
                    last = kk[e];
                    const unsigned pos = kk[e] & 0xffffu;
                    if (!((taken[pos >> 5] >> (pos & 31)) & 1u)) {
                        if (u0 == NOKEY) u0 = kk[e];
                        else if (u1 == NOKEY) u1 = kk[e];
                    }
                }
                if (u0 == NOKEY) decision = (!complete && (last >> 16) <= a.thr) ? 2 : 0;       // an unconsumed candidate <= thr may exist
                else if ((u0 >> 16) > a.thr) decision = 0;
                else if (u1 == NOKEY && !complete) decision = 2;                                  // the second best decides the ratio rule
                else {
                    const int best = (int)(u0 >> 16), lvl = a.soct[u0 & 0xffffu];
                    const int best2 = u1 != NOKEY ? (int)(u1 >> 16) : 256, lvl2 = u1 != NOKEY ? a.soct[u1 & 0xffffu] : -1;
                    decision = (lvl == lvl2 && (double)best > 0.8 * (double)best2) ? 0 : 1;       // keyframe_matcher.cpp:384-386
                }
                if (decision == 1) atomicMax(&claim[u0 & (SEARCH_CLAIMS - 1)], (round_tag << 8) | (unsigned)(31 - lane));
            }
            __syncwarp();
            bool unsafe = mine && decision == 2;
            if (mine && !unsafe) {
#pragma unroll
                for (int e = 0; e < SK; ++e) {
                    if (kk[e] == NOKEY) continue;
                    const unsigned c = claim[kk[e] & (SEARCH_CLAIMS - 1)];
                    if ((c >> 8) == round_tag && 31 - (int)(c & 0xffu) < lane) unsafe = true;
                }
            }
            const unsigned bad = __ballot_sync(0xffffffffu, unsafe);
            const int first_bad = bad ? __ffs(bad) - 1 : 32, first = __ffs(pending) - 1;
            if (first_bad == first) {
                // exact rescan of the first pending query (nothing precedes it, so it is unsafe only for this reason)
                const int qq = base + first;
                ++n_rescan;
                uint32_t d[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) d[w] = __ldg(a.qdesc + 8 * (size_t)qq + w);
                const float qx = a.qx[qq], qy = a.qy[qq], r = a.qr[qq];
                unsigned b0 = NOKEY, b1 = NOKEY;
                for (int p = range[qq].x + lane; p < range[qq].y; p += 32) {
                    if (((taken[p >> 5] >> (p & 31)) & 1u) || !in_circle(qx, qy, r, a.sx[p], a.sy[p]) || !level_ok(a, qq, a.soct[p])) continue;
                    const unsigned key = (hamming8(d, a.sdesc + 8 * (size_t)p) << 16) | (unsigned)p;
                    if (key < b0) { b1 = b0; b0 = key; } else if (key < b1) b1 = key;
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const unsigned o0 = __shfl_xor_sync(0xffffffffu, b0, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o);
                    const unsigned lo = min(b0, o0), hi = max(b0, o0);
                    b1 = min(min(b1, o1), hi);
                    b0 = lo;
                }
                if (b0 != NOKEY && (b0 >> 16) <= a.thr) {
                    const int best = (int)(b0 >> 16), lvl = a.soct[b0 & 0xffffu];
                    const int best2 = b1 != NOKEY ? (int)(b1 >> 16) : 256, lvl2 = b1 != NOKEY ? a.soct[b1 & 0xffffu] : -1;
                    if (!(lvl == lvl2 && (double)best > 0.8 * (double)best2) && lane == 0) {
                        ++count;
                        taken[(b0 & 0xffffu) >> 5] |= 1u << (b0 & 31u);
                        out_idx[qq] = a.sidx[b0 & 0xffffu];
                        out_dist[qq] = (unsigned)best;
                    }
                }
                pending &= ~(1u << first);
            } else {
                const unsigned commit = pending & ((first_bad < 32 ? (1u << first_bad) : 0u) - 1u);
                if (((commit >> lane) & 1u) && decision == 1) {
                    ++count;
                    atomicOr(&taken[(u0 & 0xffffu) >> 5], 1u << (u0 & 31u));
                    out_idx[q] = a.sidx[u0 & 0xffffu];
                    out_dist[q] = u0 >> 16;
                }
                pending &= ~commit;
            }
            __syncwarp();
        }
    }
    count = __reduce_add_sync(0xffffffffu, count);
    for (int pos = lane; pos < a.nK; pos += 32) taken_pos[pos] = (taken[pos >> 5] >> (pos & 31)) & 1u;
    if (lane == 0) { *n_matched = count; *rescans = n_rescan; }
}

// ---- medoid ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MED_THREADS)
medoid_kernel(const uint32_t *desc, const long long *offsets, int *best_out) {
    __shared__ uint4 sd[2 * MED_MAX];                       // the segment's descriptors
    __shared__ unsigned short drow[MED_THREADS / 32][MED_MAX];
    __shared__ unsigned short med[MED_MAX];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long o = offsets[s];
    const int n = (int)(offsets[s + 1] - o);
    if (n <= 0) { if (tid == 0) best_out[s] = 0; return; }
    const uint4 *src = reinterpret_cast<const uint4 *>(desc + 8 * o);
    for (int i = tid; i < 2 * n; i += MED_THREADS) sd[i] = __ldg(src + i);
    __syncthreads();
    const int kth = (n - 1) / 2;                            // static_cast<unsigned>(0.5 * (n - 1)), map_point.cpp:107
    for (int i = warp; i < n; i += MED_THREADS / 32) {
        const uint4 a0 = sd[2 * i], a1 = sd[2 * i + 1];
        for (int j = lane; j < n; j += 32) {
            const uint4 b0 = sd[2 * j], b1 = sd[2 * j + 1];
            drow[warp][j] = (unsigned short)(__popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w)
                                             + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w));
        }
        __syncwarp();
        // the (kth + 1)-th smallest = the smallest v with #(d <= v) > kth
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int c = 0;
            for (int j = lane; j < n; j += 32) c += drow[warp][j] <= mid ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c > kth) hi = mid; else lo = mid + 1;
        }
        if (lane == 0) med[i] = (unsigned short)lo;
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
        unsigned best = 0xffffffffu;
        for (int i = lane; i < n; i += 32) best = min(best, ((unsigned)med[i] << 16) | (unsigned)i);
        best = __reduce_min_sync(0xffffffffu, best);
        // `median_dist < best_median_dist` starts from 256 (map_point.cpp:99,109): a median of 256 never wins
        if (lane == 0) best_out[s] = (best >> 16) < 256u ? (int)(best & 0xffffu) : 0;
    }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_medoid(sg_ctx *ctx, const uint32_t *h_desc, const int64_t *h_offsets, int n_seg, int32_t *h_best) {
    cudaSetDevice(ctx->device);
    if (n_seg <= 0) return SG_OK;
    if (!h_desc || !h_offsets || !h_best) return fail(ctx, SG_ERR_INVALID, "null argument");
    for (int s = 0; s < n_seg; ++s) {
        const long long n = h_offsets[s + 1] - h_offsets[s];
        if (n < 0 || n > MED_MAX) return fail(ctx, SG_ERR_INVALID, "segment %d has %lld descriptors (supported: 0..%d)", s, n, MED_MAX);
    }
    const size_t total = (size_t)h_offsets[n_seg];
    Scratch sc(ctx);
    sc.want(32 * total); sc.want(8 * ((size_t)n_seg + 1)); sc.want(4 * (size_t)n_seg);
    if (int r = sc.commit()) return r;
    uint32_t *d_desc;
    long long *d_off;
    if (int r = sc.put(&d_desc, h_desc, 8 * total)) return r;
    if (int r = sc.put(&d_off, (const long long *)h_offsets, (size_t)n_seg + 1)) return r;
    int *d_best = sc.take<int>(n_seg);
    medoid_kernel<<<n_seg, MED_THREADS, 0, ctx->stream>>>(d_desc, d_off, d_best);
    SG_LAUNCH_CHECK(ctx);
    SG_CUDA(ctx, cudaMemcpyAsync(h_best, d_best, sizeof(int) * (size_t)n_seg, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

extern "C" int sg_search_candidates(sg_ctx *ctx, const float *h_kx, const float *h_ky, const int32_t *h_koct,
                                    const uint32_t *h_kdesc, int nK, const int32_t *h_order, uint8_t *h_taken,
                                    const float *h_qx, const float *h_qy, const float *h_qr, const uint32_t *h_qdesc,
                                    const int32_t *h_qlevel, int nQ, int mode, uint32_t thr, int32_t *h_idx,
                                    uint32_t *h_dist, uint32_t *n_matched) {
    cudaSetDevice(ctx->device);
    if (n_matched) *n_matched = 0;
    if (nQ <= 0) return SG_OK;
    if (!h_qx || !h_qy || !h_qr || !h_qdesc || !h_idx || !h_dist) return fail(ctx, SG_ERR_INVALID, "null argument");
    if (mode != 0 && mode != 1) return fail(ctx, SG_ERR_INVALID, "mode must be 0 or 1");
    if (nK < 0 || nK > 65535) return fail(ctx, SG_ERR_INVALID, "keypoint sets larger than 65535 are not supported");
    if (nK && (!h_kx || !h_ky || !h_koct || !h_kdesc)) return fail(ctx, SG_ERR_INVALID, "null keypoint arrays");
    // FeatureSearch index (feature_search.cpp:22-31): keypoints sorted by y with the very same std::sort call
    std::vector<int> order(nK);
    if (h_order) std::copy(h_order, h_order + nK, order.begin());
    else if (nK) sg_feature_index(h_kx, h_ky, nK, order.data());
    std::vector<float> sx(nK), sy(nK);
    std::vector<int> soct(nK);
    std::vector<uint32_t> sdesc(8 * (size_t)nK);
    std::vector<unsigned char> tk(std::max(nK, 1), 0);
    for (int p = 0; p < nK; ++p) {
        const int i = order[p];
        if (i < 0 || i >= nK) return fail(ctx, SG_ERR_INVALID, "order[%d] = %d outside the keypoint set", p, i);
        sx[p] = h_kx[i]; sy[p] = h_ky[i]; soct[p] = h_koct[i];
        std::copy(h_kdesc + 8 * (size_t)i, h_kdesc + 8 * (size_t)i + 8, sdesc.begin() + 8 * (size_t)p);
        if (mode == 1 && h_taken) tk[p] = h_taken[i] ? 1 : 0;
    }
    const size_t K = (size_t)nK, Q = (size_t)nQ;
    Scratch sc(ctx);
    sc.want(4 * K); sc.want(4 * K); sc.want(4 * K); sc.want(4 * K); sc.want(32 * K); sc.want(tk.size());
    sc.want(4 * Q); sc.want(4 * Q); sc.want(4 * Q); sc.want(32 * Q); sc.want(4 * Q);
    sc.want(16 * Q); sc.want(4 * Q); sc.want(8 * Q); sc.want(4 * Q); sc.want(4 * Q); sc.want(4); sc.want(8);
    if (int r = sc.commit()) return r;
    SearchArgs a{};
    float *d_sx, *d_sy, *d_qx, *d_qy, *d_qr;
    int *d_sidx, *d_soct, *d_ql = nullptr;
    uint32_t *d_sdesc, *d_qdesc;
    unsigned char *d_tk;
    if (int r = sc.put(&d_sx, sx.data(), K)) return r;
    if (int r = sc.put(&d_sy, sy.data(), K)) return r;
    if (int r = sc.put(&d_sidx, order.data(), K)) return r;
    if (int r = sc.put(&d_soct, soct.data(), K)) return r;
    if (int r = sc.put(&d_sdesc, sdesc.data(), 8 * K)) return r;
    if (int r = sc.put(&d_tk, tk.data(), tk.size())) return r;
    if (int r = sc.put(&d_qx, h_qx, Q)) return r;
    if (int r = sc.put(&d_qy, h_qy, Q)) return r;
    if (int r = sc.put(&d_qr, h_qr, Q)) return r;
    if (int r = sc.put(&d_qdesc, h_qdesc, 8 * Q)) return r;
    if (h_qlevel) { if (int r = sc.put(&d_ql, (const int *)h_qlevel, Q)) return r; } else sc.take<int>(Q);
    uint32_t *d_keys = sc.take<uint32_t>(4 * Q), *d_nseen = sc.take<uint32_t>(Q);
    int2 *d_range = sc.take<int2>(Q);
    int *d_idx = sc.take<int>(Q);
    uint32_t *d_dist = sc.take<uint32_t>(Q), *d_nm = sc.take<uint32_t>(1);
    unsigned long long *d_resc = sc.take<unsigned long long>(1);
    SG_CUDA(ctx, cudaMemsetAsync(d_nm, 0, 4, ctx->stream));
    SG_CUDA(ctx, cudaMemsetAsync(d_resc, 0, 8, ctx->stream));
    a.sx = d_sx; a.sy = d_sy; a.sidx = d_sidx; a.soct = d_soct; a.sdesc = d_sdesc; a.elig0 = d_tk; a.nK = nK;
    a.qx = d_qx; a.qy = d_qy; a.qr = d_qr; a.qdesc = d_qdesc; a.qlevel = d_ql; a.nQ = nQ; a.mode = mode; a.thr = thr;
    search_topk_kernel<<<(nQ + SEARCH_WARPS - 1) / SEARCH_WARPS, SEARCH_WARPS * 32, 0, ctx->stream>>>(a, d_keys, d_range, d_nseen);
    SG_LAUNCH_CHECK(ctx);
    if (mode == 0) search_resolve_independent_kernel<<<(nQ + 255) / 256, 256, 0, ctx->stream>>>(a, d_keys, d_idx, d_dist, d_nm);
    else search_resolve_sequential_kernel<<<1, 32, 0, ctx->stream>>>(a, d_keys, d_range, d_nseen, d_tk, d_idx, d_dist, d_nm, d_resc);
    SG_LAUNCH_CHECK(ctx);
    SG_CUDA(ctx, cudaMemcpyAsync(h_idx, d_idx, 4 * Q, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(h_dist, d_dist, 4 * Q, cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t nm = 0;
    SG_CUDA(ctx, cudaMemcpyAsync(&nm, d_nm, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(&ctx->rescans, d_resc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (mode == 1 && h_taken) SG_CUDA(ctx, cudaMemcpyAsync(tk.data(), d_tk, tk.size(), cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (mode == 1 && h_taken)
        for (int p = 0; p < nK; ++p) h_taken[order[p]] = tk[p];
    if (n_matched) *n_matched = nm;
    return SG_OK;
}

// matchMapPointsSim3 (keyframe_matcher.cpp:633-686): findMatchesTranformedMps (:552-631) in both directions -- two
// mode-0 candidate searches with the octave window and HAMMING_DIST_THR_HIGH -- and the agreement filter of :672-685.
static int sim3_direction(sg_ctx *ctx, const float *kx, const float *ky, const int32_t *koct, const uint32_t *kdesc, int nK,
                          const int32_t *order, const float *qx, const float *qy, const float *qr, const uint32_t *qdesc,
                          const int32_t *qlevel, int nQ, std::vector<int32_t> &match) {
    match.assign(std::max(nQ, 0), -1);
    std::vector<int> live;
    for (int q = 0; q < nQ; ++q)
        if (qr[q] >= 0.f) live.push_back(q);          // r < 0: the keypoint issues no query (:566-590)
    if (live.empty() || nK == 0) return SG_OK;
    const size_t L = live.size();
    std::vector<float> x(L), y(L), r(L);
    std::vector<int32_t> lvl(L), idx(L);
    std::vector<uint32_t> desc(8 * L), dist(L);
    for (size_t i = 0; i < L; ++i) {
        const int q = live[i];
        x[i] = qx[q]; y[i] = qy[q]; r[i] = qr[q]; lvl[i] = qlevel[q];
        std::copy(qdesc + 8 * (size_t)q, qdesc + 8 * (size_t)q + 8, desc.begin() + 8 * i);
    }
    if (int rc = sg_search_candidates(ctx, kx, ky, koct, kdesc, nK, order, nullptr, x.data(), y.data(), r.data(), desc.data(),
                                      lvl.data(), (int)L, 0, 100u, idx.data(), dist.data(), nullptr))
        return rc;
    for (size_t i = 0; i < L; ++i) match[live[i]] = idx[i];
    return SG_OK;
}

extern "C" int sg_match_sim3(sg_ctx *ctx, const float *h_x1, const float *h_y1, const int32_t *h_oct1, const uint32_t *h_desc1,
                             int n1, const int32_t *h_order1, const float *h_x2, const float *h_y2, const int32_t *h_oct2,
                             const uint32_t *h_desc2, int n2, const int32_t *h_order2, const float *h_q12x, const float *h_q12y,
                             const float *h_q12r, const uint32_t *h_q12desc, const int32_t *h_q12level, const float *h_q21x,
                             const float *h_q21y, const float *h_q21r, const uint32_t *h_q21desc, const int32_t *h_q21level,
                             int32_t *h_pairs, uint32_t *n_pairs) {
    if (n_pairs) *n_pairs = 0;
    if (n1 < 0 || n2 < 0) return fail(ctx, SG_ERR_INVALID, "negative keypoint count");
    if (n1 == 0 || n2 == 0) return SG_OK;
    if (!h_q12x || !h_q12y || !h_q12r || !h_q12desc || !h_q12level || !h_q21x || !h_q21y || !h_q21r || !h_q21desc || !h_q21level
        || !h_pairs)
        return fail(ctx, SG_ERR_INVALID, "null argument");
    std::vector<int32_t> m12, m21;
    if (int rc = sim3_direction(ctx, h_x2, h_y2, h_oct2, h_desc2, n2, h_order2, h_q12x, h_q12y, h_q12r, h_q12desc, h_q12level, n1, m12))
        return rc;
    if (int rc = sim3_direction(ctx, h_x1, h_y1, h_oct1, h_desc1, n1, h_order1, h_q21x, h_q21y, h_q21r, h_q21desc, h_q21level, n2, m21))
        return rc;
    uint32_t n = 0;
    for (int i = 0; i < n1; ++i) {
        const int j = m12[i];
        if (j < 0) continue;
        if (m21[j] == i) { h_pairs[2 * n] = i; h_pairs[2 * n + 1] = j; ++n; }   // 1 -> 2 and 2 -> 1 agree (:680)
    }
    if (n_pairs) *n_pairs = n;
    return SG_OK;
}

extern "C" int sg_feature_index(const float *h_x, const float *h_y, int n, int32_t *h_order) {
    struct Node { float x, y; int idx; };
    std::vector<Node> v(std::max(n, 0));
    for (int i = 0; i < n; ++i) v[i] = Node{h_x[i], h_y[i], i};
    std::sort(v.begin(), v.end(), [](const Node &a, const Node &b) { return a.y < b.y; });
    for (int i = 0; i < n; ++i) h_order[i] = v[i].idx;
    return SG_OK;
}
