// Brute-force Hamming matching with the sequential semantics of matchForLoopClosures
// (keyframe_matcher.cpp:50-158, single-BoW-node case): for every A feature in index order, best
// and second-best distance over the B features not yet consumed, reject when thr < best
// (:115) or ratio * second < best (:120), consume the B feature (:128), finally keep only matches
// whose delta-angle falls in the three most populated 30-degree bins
// (openvslam/match_angle_checker.h:61-134).  Distances are compute_descriptor_distance_32
// (openvslam/match_base.h:18-39) = popcount of the XOR of 256 bits.
//
// GPU formulation (exact, not approximate)
//   hamming_topk_kernel : the O(nA*nB) part.  One thread owns one A descriptor in registers and
//     streams all B descriptors from a shared-memory tile (broadcast 128-bit reads), 8 XOR + 8 POPC
//     per pair, and keeps the K = 4 smallest (distance, index) keys with distance <= C, where C is
//     the largest "second best" that can still change a ratio decision (host, match_cutoff).
//   match_resolve_kernel: the order-dependent part.  One warp per keyframe pair walks the A rows (32 at a time, optimistically: see the kernel)
//     in order; best / second are the first two unconsumed entries of a row's list.  When the list
//     cannot decide (entries consumed and the list was truncated) the warp rescans the whole row
//     exactly.  Then the angle histogram filter, with libstdc++'s std::sort order of the 30 bins
//     restated (sort30) so that ties between bins resolve as in the reference.
#include <algorithm>
#include "ctx.h"

namespace sg {

constexpr int TOPK = 4;           // list length of the throughput path (and of run_topk_lists)
constexpr int TOPK_WIDE = 16;     // launches with few pairs: longer lists, (almost) no exact rescans on clustered descriptors
constexpr int MT_THREADS = 256;
constexpr int B_CHUNK = 1024;        // B descriptors per shared-memory tile (32 KB)
constexpr unsigned EMPTY_KEY = 0xffffffffu;
constexpr int RES_WARPS = 4;
constexpr int CLAIM_SLOTS = 2048;    // claim table of the optimistic resolve (power of two)

// ---- libstdc++ std::sort (introsort, _S_threshold 16) restated for 30 bin indices -----------------
// comp(a, b) == sizes[a] > sizes[b]  (match_angle_checker.h:129-132).  Equal sizes make the result
// depend on the algorithm, so it is restated step by step: median-of-three to first, unguarded
// Hoare partition, recursion on the right part, heap sort when the depth limit is hit, final
// insertion sort (first 16 guarded, rest unguarded).
struct Sort30 {
    unsigned v[30];
    const unsigned *sz;
    __host__ __device__ bool comp(unsigned a, unsigned b) const { return sz[a] > sz[b]; }
    __host__ __device__ void swp(int i, int j) { const unsigned t = v[i]; v[i] = v[j]; v[j] = t; }

    __host__ __device__ void median_to_first(int result, int a, int b, int c) {
        if (comp(v[a], v[b])) {
            if (comp(v[b], v[c])) swp(result, b);
            else if (comp(v[a], v[c])) swp(result, c);
            else swp(result, a);
        } else if (comp(v[a], v[c])) swp(result, a);
        else if (comp(v[b], v[c])) swp(result, c);
        else swp(result, b);
    }
    __host__ __device__ int partition(int first, int last, int pivot) {
        while (true) {
            while (comp(v[first], v[pivot])) ++first;
            --last;
            while (comp(v[pivot], v[last])) --last;
            if (!(first < last)) return first;
            swp(first, last);
            ++first;
        }
    }
    // heap helpers (std::__adjust_heap / __push_heap / __pop_heap / __make_heap) on [first, first+len)
    __host__ __device__ void adjust_heap(int first, int hole, int len, unsigned value) {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (comp(v[first + child], v[first + child - 1])) --child;
            v[first + hole] = v[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            v[first + hole] = v[first + child - 1];
            hole = child - 1;
        }
        int parent = (hole - 1) / 2;
        while (hole > top && comp(v[first + parent], value)) {
            v[first + hole] = v[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        v[first + hole] = value;
    }
    __host__ __device__ void heap_sort(int first, int last) {   // std::__partial_sort(first, last, last)
        const int len = last - first;
        if (len >= 2)
            for (int parent = (len - 2) / 2;; --parent) {
                adjust_heap(first, parent, len, v[first + parent]);
                if (parent == 0) break;
            }
        for (int l = last; l - first > 1;) {
            --l;
            const unsigned value = v[l];
            v[l] = v[first];
            adjust_heap(first, 0, l - first, value);
        }
    }
    __host__ __device__ void introsort_loop(int first0, int last0, int depth0) {
        // the recursion on [cut, last) touches a range disjoint from the loop's [first, cut), so an explicit
        // stack visiting the parts in any order gives the same result
        int sf[32], sl[32], sd[32], sp = 0;
        sf[0] = first0; sl[0] = last0; sd[0] = depth0; sp = 1;
        while (sp) {
            --sp;
            int first = sf[sp], last = sl[sp], depth = sd[sp];
            while (last - first > 16) {
                if (depth == 0) { heap_sort(first, last); break; }
                --depth;
                const int mid = first + (last - first) / 2;
                median_to_first(first, first + 1, mid, last - 1);
                const int cut = partition(first + 1, last, first);
                sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;
                last = cut;
            }
        }
    }
    __host__ __device__ void linear_insert(int last) {
        const unsigned val = v[last];
        int next = last - 1;
        while (comp(val, v[next])) { v[last] = v[next]; last = next; --next; }
        v[last] = val;
    }
    __host__ __device__ void insertion_sort(int first, int last) {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (comp(v[i], v[first])) {
                const unsigned val = v[i];
                for (int k = i; k > first; --k) v[k] = v[k - 1];
                v[first] = val;
            } else linear_insert(i);
        }
    }
    __host__ __device__ void run(const unsigned *sizes, int depth_limit) {
        sz = sizes;
        for (int i = 0; i < 30; ++i) v[i] = i;
        introsort_loop(0, 30, depth_limit);
        insertion_sort(0, 16);
        for (int i = 16; i < 30; ++i) linear_insert(i);
    }
};

// angle_checker::append_delta_angle (match_angle_checker.h:72-83): float -> double compare/add -> float
__host__ __device__ inline int angle_bin(float delta) {
    if (delta < 0.0) delta = (float)((double)delta + 360.0);
    if (360.0 <= delta) delta = (float)((double)delta - 360.0);
#ifdef __CUDA_ARCH__
    return __float2int_rn(__fmul_rn(delta, 1.0f / 30));
#else
    return (int)lrintf(delta * (1.0f / 30));
#endif
}

// ---- distance + top-K kernel -----------------------------------------------------------------------
struct MatchArgs {
    const uint32_t *desc;      // database descriptors
    const float *angle;
    const long long *offsets;  // set offsets
    const int *pairs;          // {setA, setB} per pair
    int row_stride;            // rows reserved per pair in the top-K scratch
    int match_stride;          // ints per pair in the match output
    unsigned cutoff;           // C
    unsigned thr;
    float ratio;
    int ratio_is_double;
    int check_orientation;
    int last_wins;             // ties between equal distances resolve to the LARGER B index (matchForTriangulationDBoW :231)
    int splits;                // the B set is cut into `splits` ranges (grid z); > 1 only when there are few pairs
    long long split_rows;      // scratch rows per split (pairs in the launch x row_stride)
};

template <int K>
__device__ __forceinline__ void topk_insert(unsigned (&t)[K], unsigned key) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const unsigned lo = min(t[i], key);
        key = max(t[i], key);
        t[i] = lo;
    }
}

// popcount of the 256-bit XOR with carry-save adders: three full adders (two LOP3 each) compress seven of the
// eight XOR words into one word of weight 1, one of weight 2 and one of weight 4, so 4 POPC (quarter-rate pipe)
// replace 8; the result is the same integer as compute_descriptor_distance_32 (openvslam/match_base.h:18-39).
__device__ __forceinline__ void csa(unsigned a, unsigned b, unsigned c, unsigned &sum, unsigned &carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a | b));
}
__device__ __forceinline__ unsigned hamming256_csa(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    unsigned s1, c1, s2, c2, s3, c3, s4, c4;
    csa(a0.x ^ b0.x, a0.y ^ b0.y, a0.z ^ b0.z, s1, c1);
    csa(a0.w ^ b0.w, a1.x ^ b1.x, a1.y ^ b1.y, s2, c2);
    csa(s1, s2, a1.z ^ b1.z, s3, c3);
    csa(c1, c2, c3, s4, c4);
    return __popc(s3) + __popc(a1.w ^ b1.w) + 2u * __popc(s4) + 4u * __popc(c4);
}

template <int K>
__global__ void __launch_bounds__(MT_THREADS)
hamming_topk_kernel(const MatchArgs a, uint32_t *topk, uint32_t *nseen_out) {
    __shared__ uint4 Bs[B_CHUNK * 2];
    const int tid = threadIdx.x, p = blockIdx.y;
    const int sa = a.pairs[2 * p], sb = a.pairs[2 * p + 1];
    const long long oa = a.offsets[sa], ob = a.offsets[sb];
    const int nA = (int)(a.offsets[sa + 1] - oa), nB = (int)(a.offsets[sb + 1] - ob);
    if ((int)(blockIdx.x * MT_THREADS) >= nA) return;
    const int row = blockIdx.x * MT_THREADS + tid;
    const bool live = row < nA;

    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (live) {
        const uint4 *pa = reinterpret_cast<const uint4 *>(a.desc + 8 * (oa + row));
        a0 = __ldg(pa); a1 = __ldg(pa + 1);
    }
    unsigned t[K];
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = EMPTY_KEY;
    unsigned nseen = 0;
    const unsigned C = a.cutoff;
    const uint4 *pb = reinterpret_cast<const uint4 *>(a.desc + 8 * ob);
    // few pairs in the launch: the B rows are shared out over grid z so that the chip is filled; the partial lists
    // are merged by topk_merge_kernel
    const int per_split = (nB + a.splits - 1) / a.splits;
    const int jb = min(nB, (int)blockIdx.z * per_split), je = min(nB, jb + per_split);

    for (int j0 = jb; j0 < je; j0 += B_CHUNK) {
        const int cn = min(B_CHUNK, je - j0);
        __syncthreads();
        for (int i = tid; i < 2 * cn; i += MT_THREADS) Bs[i] = __ldg(pb + 2 * (size_t)j0 + i);
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int j = 0; j < cn; ++j) {
                const uint4 b0 = Bs[2 * j], b1 = Bs[2 * j + 1];
                const unsigned d = hamming256_csa(a0, a1, b0, b1);
                if (d <= C) {
                    ++nseen;
                    const unsigned key = (d << 16) | (a.last_wins ? 0xffffu - (unsigned)(j0 + j) : (unsigned)(j0 + j));
                    if (key < t[K - 1]) topk_insert(t, key);
                }
            }
        }
    }
    if (live) {
        const size_t r = (size_t)blockIdx.z * a.split_rows + (size_t)p * a.row_stride + row;
#pragma unroll
        for (int q = 0; q < K / 4; ++q)
            reinterpret_cast<uint4 *>(topk)[r * (K / 4) + q] = make_uint4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
        nseen_out[r] = nseen;
    }
}

// Merge of the per-split lists into split 0: the 4 smallest keys (keys are unique: they carry the B index) and the
// sum of the counts.  Rows no split wrote (row >= nA) hold stale keys that the resolve kernel never reads.
template <int K>
__global__ void topk_merge_kernel(uint32_t *topk, uint32_t *nseen, int splits, long long split_rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= split_rows) return;
    unsigned t[K];
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const uint4 v = reinterpret_cast<const uint4 *>(topk)[i * (K / 4) + q];
        t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
    }
    unsigned n = nseen[i];
    for (int sp = 1; sp < splits; ++sp) {
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            const uint4 v = reinterpret_cast<const uint4 *>(topk)[(sp * split_rows + i) * (K / 4) + q];
            if (v.x < t[K - 1]) topk_insert(t, v.x);
            if (v.y < t[K - 1]) topk_insert(t, v.y);
            if (v.z < t[K - 1]) topk_insert(t, v.z);
            if (v.w < t[K - 1]) topk_insert(t, v.w);
        }
        n += nseen[sp * split_rows + i];
    }
#pragma unroll
    for (int q = 0; q < K / 4; ++q)
        reinterpret_cast<uint4 *>(topk)[i * (K / 4) + q] = make_uint4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
    nseen[i] = n;
}

// How many B ranges a launch of `ctas` CTAs (row tiles x pairs) should use: 1 when the grid already fills the chip.
static int pick_splits(int ctas, int max_set) {
    if (ctas >= 2 * 148) return 1;
    const int want = (4 * 148 + ctas - 1) / std::max(ctas, 1);
    return std::max(1, std::min(std::min(want, 8), max_set / 128));
}

template <int K>
static int launch_topk(sg_ctx *ctx, MatchArgs &a, int tiles, int np, int splits, uint32_t *d_topk, uint32_t *d_nseen) {
    a.splits = splits;
    a.split_rows = (long long)np * a.row_stride;
    hamming_topk_kernel<K><<<dim3(tiles, np, splits), MT_THREADS, 0, ctx->stream>>>(a, d_topk, d_nseen);
    SG_LAUNCH_CHECK(ctx);
    if (splits > 1) {
        topk_merge_kernel<K><<<(unsigned)((a.split_rows + 255) / 256), 256, 0, ctx->stream>>>(d_topk, d_nseen, splits, a.split_rows);
        SG_LAUNCH_CHECK(ctx);
    }
    return SG_OK;
}

// ---- sequential resolve + angle filter -------------------------------------------------------------
__device__ __forceinline__ bool ratio_rejects(float ratio, unsigned second, unsigned best, int is_double) {
    // keyframe_matcher.cpp:120  `ratio * second < static_cast<float>(best)`
    if (is_double) return (double)ratio * (double)second < (double)(float)best;
    return __fmul_rn(ratio, (float)second) < (float)best;
}

template <int K>
__global__ void __launch_bounds__(RES_WARPS * 32)
match_resolve_kernel(const MatchArgs a, int n_pairs, const uint32_t *topk, const uint32_t *nseen_in,
                     int taken_words, int *matches, uint32_t *n_matches, unsigned long long *rescans) {
    extern __shared__ uint32_t rsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * RES_WARPS + warp;
    if (p >= n_pairs) return;
    uint32_t *taken = rsm + warp * (taken_words + 32 + CLAIM_SLOTS);
    uint32_t *hist = taken + taken_words;   // 32 words, 30 used
    uint32_t *claim = hist + 32;            // [CLAIM_SLOTS], slot = B index mod CLAIM_SLOTS: round tag << 8 | 31 - lane of the earliest
                                            // claiming row (a slot shared by two indices can only make a row look unsafe: conservative)

    const int sa = a.pairs[2 * p], sb = a.pairs[2 * p + 1];
    const long long oa = a.offsets[sa], ob = a.offsets[sb];
    const int nA = (int)(a.offsets[sa + 1] - oa), nB = (int)(a.offsets[sb + 1] - ob);
    const uint32_t *dA = a.desc + 8 * oa, *dB = a.desc + 8 * ob;
    const float *angA = a.angle + oa, *angB = a.angle + ob;
    int *mrow = matches + (size_t)p * a.match_stride;

    for (int i = lane; i < taken_words; i += 32) taken[i] = 0;
    for (int i = lane; i < CLAIM_SLOTS; i += 32) claim[i] = 0;
    for (int i = nA + lane; i < a.match_stride; i += 32) mrow[i] = -1;   // padding behind the A set
    hist[lane] = 0;
    __syncwarp();

    unsigned count = 0, n_rescan = 0, round_tag = 0;
    // the match row is also the record the angle filter re-reads, so it always exists: the caller's
    // buffer or the context's scratch (run_match)
    for (int base = 0; base < nA; base += 32) {
        const int i = base + lane;
        unsigned kk[K];
#pragma unroll
        for (int e = 0; e < K; ++e) kk[e] = EMPTY_KEY;
        unsigned ns = 0;
        if (i < nA) {
            const size_t r = (size_t)p * a.row_stride + i;
            ns = nseen_in[r];
            if (ns) {
#pragma unroll
                for (int q = 0; q < K / 4; ++q) {
                    const uint4 v = reinterpret_cast<const uint4 *>(topk)[r * (K / 4) + q];
                    kk[4 * q] = v.x; kk[4 * q + 1] = v.y; kk[4 * q + 2] = v.z; kk[4 * q + 3] = v.w;
                }
            }
            mrow[i] = -1;
        }
        // Optimistic batch: every lane evaluates its own row against the current `taken` bits; accepted rows
        // claim their B index; a row is SAFE when no earlier pending row of the batch claimed one of its keys.
        // The safe prefix is committed at once (its tentative decisions are exactly the sequential ones), the
        // first unsafe row is re-evaluated in the next round (it is then first and therefore safe) or, when its
        // truncated list cannot decide, rescanned exactly by the whole warp.
        unsigned pending = __ballot_sync(0xffffffffu, ns > 0);
        const bool complete = ns <= (unsigned)K;
        while (pending) {
            ++round_tag;
            int decision = 0;   // 0 reject, 1 accept u0, 2 rescan
            unsigned best_idx = 0;
            const bool mine = (pending >> lane) & 1u;
            if (mine) {
                unsigned u0 = EMPTY_KEY, u1 = EMPTY_KEY, last = EMPTY_KEY;
#pragma unroll
                for (int e = 0; e < K; ++e) {
                    if (kk[e] == EMPTY_KEY) continue;
                    last = kk[e];
                    const unsigned idx = kk[e] & 0xffffu;
                    if (!((taken[idx >> 5] >> (idx & 31)) & 1u)) {
                        if (u0 == EMPTY_KEY) u0 = kk[e];
                        else if (u1 == EMPTY_KEY) u1 = kk[e];
                    }
                }
                const unsigned last_d = last >> 16;
                if (u0 == EMPTY_KEY) {
                    decision = (complete || last_d > a.thr) ? 0 : 2;
                } else {
                    const unsigned best = u0 >> 16;
                    if (best > a.thr) decision = 0;
                    else if (u1 != EMPTY_KEY) decision = ratio_rejects(a.ratio, u1 >> 16, best, a.ratio_is_double) ? 0 : 1;
                    else if (complete) decision = ratio_rejects(a.ratio, 256u, best, a.ratio_is_double) ? 0 : 1;
                    else decision = ratio_rejects(a.ratio, last_d, best, a.ratio_is_double) ? 2 : 1;
                }
                best_idx = u0 & 0xffffu;
                if (decision == 1) atomicMax(&claim[best_idx & (CLAIM_SLOTS - 1)], (round_tag << 8) | (unsigned)(31 - lane));   // earliest row wins
            }
            __syncwarp();
            bool unsafe = mine && decision == 2;
            if (mine && !unsafe) {
#pragma unroll
                for (int e = 0; e < K; ++e) {
                    if (kk[e] == EMPTY_KEY) continue;
                    const unsigned c = claim[kk[e] & (CLAIM_SLOTS - 1)];
                    if ((c >> 8) == round_tag && 31 - (int)(c & 0xffu) < lane) unsafe = true;   // claimed by an earlier row
                }
            }
            const unsigned bad = __ballot_sync(0xffffffffu, unsafe);
            const int first_bad = bad ? __ffs(bad) - 1 : 32;
            const int first = __ffs(pending) - 1;
            if (first_bad == first) {
                // the first pending row is unsafe only when it needs the exact full-row scan (nothing precedes it)
                const int row = base + first;
                ++n_rescan;
                uint32_t ar[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) ar[w] = __ldg(dA + 8 * (size_t)row + w);
                unsigned bkey = (256u << 16) | 0xffffu, sec = 256u;
                // four independent 32-row groups per iteration: the loads of all four are in flight together (the scan is
                // a chain of global-memory round trips otherwise); best / second do not depend on the visiting order
                constexpr int RU = 4;
                for (int j0 = lane; j0 < nB; j0 += 32 * RU) {
                    uint4 b0[RU], b1[RU];
                    bool ok[RU];
#pragma unroll
                    for (int u = 0; u < RU; ++u) {
                        const int j = j0 + 32 * u;
                        const int jc = min(j, nB - 1);               // unconditional (clamped) loads: no branch between them
                        ok[u] = j < nB && !((taken[jc >> 5] >> (jc & 31)) & 1u);
                        b0[u] = __ldg(reinterpret_cast<const uint4 *>(dB + 8 * (size_t)jc));
                        b1[u] = __ldg(reinterpret_cast<const uint4 *>(dB + 8 * (size_t)jc) + 1);
                    }
#pragma unroll
                    for (int u = 0; u < RU; ++u) {
                        if (!ok[u]) continue;
                        const unsigned d = __popc(ar[0] ^ b0[u].x) + __popc(ar[1] ^ b0[u].y) + __popc(ar[2] ^ b0[u].z) + __popc(ar[3] ^ b0[u].w)
                                           + __popc(ar[4] ^ b1[u].x) + __popc(ar[5] ^ b1[u].y) + __popc(ar[6] ^ b1[u].z) + __popc(ar[7] ^ b1[u].w);
                        const unsigned key = (d << 16) | (unsigned)(j0 + 32 * u);
                        if (key < bkey) { sec = bkey >> 16; bkey = key; }
                        else if (d < sec) sec = d;
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const unsigned ob_ = __shfl_xor_sync(0xffffffffu, bkey, o);
                    const unsigned os_ = __shfl_xor_sync(0xffffffffu, sec, o);
                    sec = min(min(sec, os_), max(bkey >> 16, ob_ >> 16));
                    bkey = min(bkey, ob_);
                }
                const unsigned best = bkey >> 16;
                if (!(a.thr < best || ratio_rejects(a.ratio, sec, best, a.ratio_is_double))) {
                    if (lane == 0) {
                        ++count;
                        taken[(bkey & 0xffffu) >> 5] |= 1u << (bkey & 31u);
                        mrow[row] = (int)(bkey & 0xffffu);
                    }
                }
                pending &= ~(1u << first);
            } else {
                // commit the safe prefix: pending rows before the first unsafe one
                const unsigned commit = pending & ((first_bad < 32 ? (1u << first_bad) : 0u) - 1u);
                if (((commit >> lane) & 1u) && decision == 1) {
                    ++count;
                    atomicOr(&taken[best_idx >> 5], 1u << (best_idx & 31));
                    mrow[base + lane] = (int)best_idx;
                }
                pending &= ~commit;
            }
            __syncwarp();
        }
    }
    count = __reduce_add_sync(0xffffffffu, count);
    // ---- angle histogram filter ------------------------------------------------------------------------
    if (a.check_orientation) {
        // delta-angle histogram of the accepted matches, after the sequential walk: the angle loads are global
        // memory round trips that must not sit inside the row-by-row dependency chain
        __syncwarp();
        for (int i = lane; i < nA; i += 32) {
            const int m = mrow[i];
            if (m >= 0) atomicAdd(&hist[angle_bin(angA[i] - angB[m])], 1u);
        }
        __syncwarp();
        unsigned valid = 0;
        if (lane == 0) {
            Sort30 s;
            s.run(hist, 8);   // depth limit 2 * floor(log2(30))
            valid = (1u << s.v[0]) | (1u << s.v[1]) | (1u << s.v[2]);
        }
        valid = __shfl_sync(0xffffffffu, valid, 0);
        unsigned removed = 0;
        for (int i = lane; i < nA; i += 32) {
            const int m = mrow[i];
            if (m >= 0) {
                const int bin = angle_bin(angA[i] - angB[m]);
                if (!((valid >> bin) & 1u)) { mrow[i] = -1; ++removed; }
            }
        }
        removed = __reduce_add_sync(0xffffffffu, removed);
        count -= removed;
    }
    if (lane == 0) {
        n_matches[p] = count;
        if (n_rescan) atomicAdd(rescans, (unsigned long long)n_rescan);
    }
}

// Largest second-best distance that can still flip the ratio test for some best <= thr.
static unsigned match_cutoff(const sg_match_params &mp) {
    unsigned c = mp.thr;
    while (c < 256) {
        const unsigned s = c + 1;
        const bool rej = mp.ratio_is_double ? ((double)mp.ratio * (double)s < (double)(float)mp.thr)
                                            : (mp.ratio * (float)s < (float)mp.thr);
        if (!rej) break;
        c = s;
    }
    return std::min(c, 256u);
}

// ---- simple kernels: pairwise distances, POPC micro-benchmark ---------------------------------------
__global__ void hamming_pairs_kernel(const uint32_t *a, const uint32_t *b, int n, uint32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *pa = reinterpret_cast<const uint4 *>(a) + 2 * (size_t)i, *pb = reinterpret_cast<const uint4 *>(b) + 2 * (size_t)i;
    const uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
    out[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w)
             + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void popc_bench_kernel(unsigned *out, int iters, unsigned seed) {
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u + blockIdx.x;
    unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] += __popc(x[i]);   // 8 independent POPC per iteration
            x[i] ^= acc[i] + it;      // keeps the compiler from hoisting; one LOP3/IADD per POPC
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 0xdeadbeefu) out[0] = s;
}

// Pairs per launch: bounds the top-K (and match-row) scratch to ~256 MB.
int match_chunk_pairs(const sg_db *db, bool own_matches) {
    const size_t per_pair = (size_t)std::max(db->max_set, 1) * (TOPK * 4 + 4 + (own_matches ? 4 : 0));
    // grid.y of hamming_topk_kernel is the pair index: at most 65535 pairs per launch whatever the scratch allows
    return (int)std::min<size_t>(65535, std::max<size_t>(1, ((size_t)256 << 20) / per_pair));
}

// Top-K scratch: `rows` rows of `K` keys + one count per row, from the stream-ordered pool.
static int ensure_topk(sg_ctx *ctx, size_t rows, int K) {
    if (rows <= ctx->topk_rows && rows * K <= ctx->topk_words) return SG_OK;
    rows = std::max(rows, ctx->topk_rows);
    const size_t words = std::max(rows * K, ctx->topk_words);
    if (ctx->d_topk) cudaFreeAsync(ctx->d_topk, ctx->main_stream);
    if (ctx->d_nseen) cudaFreeAsync(ctx->d_nseen, ctx->main_stream);
    ctx->d_topk = nullptr; ctx->d_nseen = nullptr; ctx->topk_rows = 0; ctx->topk_words = 0;
    SG_CUDA(ctx, cudaMallocFromPoolAsync((void **)&ctx->d_topk, words * 4, ctx->pool, ctx->main_stream));
    SG_CUDA(ctx, cudaMallocFromPoolAsync((void **)&ctx->d_nseen, rows * 4, ctx->pool, ctx->main_stream));
    ctx->topk_rows = rows; ctx->topk_words = words;
    return SG_OK;
}

template <int K>
static int run_match_k(sg_ctx *ctx, const sg_db *db, const int *d_pairs, int n_pairs, const sg_match_params &mp,
                       int *d_matches, int match_stride, uint32_t *d_n_matches, int chunk, int tiles, int splits) {
    const int stride = db->max_set;
    MatchArgs a{};
    a.desc = db->d_desc; a.angle = db->d_angle; a.offsets = db->d_offsets;
    a.cutoff = match_cutoff(mp); a.thr = mp.thr; a.ratio = mp.ratio;
    a.ratio_is_double = mp.ratio_is_double; a.check_orientation = mp.check_orientation;
    const int taken_words = (db->max_set + 31) / 32;
    const size_t rsmem = (size_t)RES_WARPS * (taken_words + 32 + CLAIM_SLOTS) * 4;
    if (rsmem > 48 * 1024)
        SG_CUDA(ctx, cudaFuncSetAttribute(match_resolve_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    for (int p0 = 0; p0 < n_pairs; p0 += chunk) {
        const int np = std::min(chunk, n_pairs - p0);
        if (p0 == 0) mark(ctx, EV_MATCH0, true);
        a.pairs = d_pairs + 2 * (size_t)p0;
        a.row_stride = stride;
        int sp = np == chunk ? splits : pick_splits(tiles * np, db->max_set);
        sp = (int)std::max<size_t>(1, std::min<size_t>(sp, ctx->topk_rows / ((size_t)np * stride)));   // never beyond the scratch
        if (int r = launch_topk<K>(ctx, a, tiles, np, sp, ctx->d_topk, ctx->d_nseen)) return r;
        if (p0 == 0) mark(ctx, EV_TOPK1);
        a.match_stride = d_matches ? match_stride : stride;
        int *mout = d_matches ? d_matches + (size_t)p0 * match_stride : ctx->d_matches;
        match_resolve_kernel<K><<<(np + RES_WARPS - 1) / RES_WARPS, RES_WARPS * 32, rsmem, ctx->stream>>>(
            a, np, ctx->d_topk, ctx->d_nseen, taken_words, mout, d_n_matches + p0, ctx->d_rescans);
        SG_LAUNCH_CHECK(ctx);
        if (p0 == 0) mark(ctx, EV_RESOLVE1);
    }
    return SG_OK;
}

int run_match(sg_ctx *ctx, const sg_db *db, const int *d_pairs, int n_pairs, const sg_match_params &mp,
              int *d_matches, int match_stride, uint32_t *d_n_matches, bool reset_rescans) {
    if (n_pairs <= 0) return SG_OK;
    if (db->max_set > 65535) return fail(ctx, SG_ERR_INVALID, "descriptor sets larger than 65535 features are not supported");
    if (mp.thr > 256) return fail(ctx, SG_ERR_INVALID, "thr must be <= 256");
    const int stride = db->max_set;
    if (d_matches && match_stride < stride) return fail(ctx, SG_ERR_INVALID, "match_stride smaller than the largest set");
    if (stride == 0) {   // every set is empty: zero matches for every pair, no launch
        SG_CUDA(ctx, cudaMemsetAsync(d_n_matches, 0, sizeof(uint32_t) * (size_t)n_pairs, ctx->stream));
        if (d_matches && match_stride > 0)
            SG_CUDA(ctx, cudaMemsetAsync(d_matches, 0xff, sizeof(int) * (size_t)n_pairs * match_stride, ctx->stream));
        return SG_OK;
    }
    const int chunk = std::min(n_pairs, match_chunk_pairs(db, d_matches == nullptr));
    const int tiles = (stride + MT_THREADS - 1) / MT_THREADS;
    const int splits = pick_splits(tiles * chunk, db->max_set);
    // Few pairs (single-call matchers, stereo stream): latency matters and the grid is small anyway, so the lists are
    // 16 entries long -- clusters of near-identical descriptors (repeated structure) then hardly ever exhaust a list,
    // which would cost an exact rescan by a single warp.  Many pairs: 4 entries, the throughput configuration.
    const bool wide = tiles * chunk < 2 * 148;
    const size_t need = (size_t)chunk * stride;
    if (int r = ensure_topk(ctx, need * splits, wide ? TOPK_WIDE : TOPK)) return r;
    if (!d_matches) {
        size_t cap = ctx->matches_cap;
        int r = grow(ctx, (void **)&ctx->d_matches, &cap, need, sizeof(int));
        ctx->matches_cap = cap;
        if (r) return r;
    }
    if (reset_rescans) SG_CUDA(ctx, cudaMemsetAsync(ctx->d_rescans, 0, sizeof(unsigned long long), ctx->stream));
    return wide ? run_match_k<TOPK_WIDE>(ctx, db, d_pairs, n_pairs, mp, d_matches, match_stride, d_n_matches, chunk, tiles, splits)
                : run_match_k<TOPK>(ctx, db, d_pairs, n_pairs, mp, d_matches, match_stride, d_n_matches, chunk, tiles, splits);
}

// ---- candidate lists for matchForTriangulationDBoW (keyframe_matcher.cpp:160-293) ----------------------------------
// That matcher keeps, per kf1 feature, the LAST kf2 feature (in node order) whose distance is <= 50 and <= the best so
// far AND that passes the fp64 epipolar test (:231-242).  The Hamming part runs here: per row the 4 smallest keys
// (distance << 16 | 0xffff - index), i.e. ascending distance, ties by descending index -- the order in which the host
// has to try the epipolar test.  A row whose list is truncated and exhausted is rescanned exactly by row_scan_kernel.
int run_topk_lists(sg_ctx *ctx, const sg_db *db, const int *d_pairs, int n_pairs, unsigned thr, uint32_t **d_topk,
                   uint32_t **d_nseen, int *row_stride) {
    if (db->max_set > 65535) return fail(ctx, SG_ERR_INVALID, "descriptor sets larger than 65535 features are not supported");
    const int stride = std::max(db->max_set, 1);
    const int tiles = (stride + MT_THREADS - 1) / MT_THREADS;
    const int splits = pick_splits(tiles * n_pairs, db->max_set);
    if (int r = ensure_topk(ctx, (size_t)n_pairs * stride * splits, TOPK)) return r;
    MatchArgs a{};
    a.desc = db->d_desc; a.angle = db->d_angle; a.offsets = db->d_offsets; a.pairs = d_pairs;
    a.cutoff = thr; a.thr = thr; a.row_stride = stride; a.last_wins = 1;
    if (int r = launch_topk<TOPK>(ctx, a, tiles, n_pairs, splits, ctx->d_topk, ctx->d_nseen)) return r;
    *d_topk = ctx->d_topk; *d_nseen = ctx->d_nseen; *row_stride = stride;
    return SG_OK;
}

// All B indices of a set within `thr` of one A descriptor: rows[r] = {setA row (absolute), setB}; out[r][..] unsorted.
__global__ void __launch_bounds__(128)
row_scan_kernel(const uint32_t *desc, const long long *offsets, const int *rows, int n_rows, unsigned thr, int out_stride,
                uint32_t *out, uint32_t *out_count) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const long long rowA = rows[2 * r];
    const int sb = rows[2 * r + 1];
    const long long ob = offsets[sb];
    const int nB = (int)(offsets[sb + 1] - ob);
    const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(desc + 8 * rowA)), a1 = __ldg(reinterpret_cast<const uint4 *>(desc + 8 * rowA) + 1);
    unsigned n = 0;
    for (int j0 = 0; j0 < nB; j0 += 32) {
        const int j = j0 + lane;
        unsigned d = 0xffffffffu;
        if (j < nB) {
            const uint4 *pb = reinterpret_cast<const uint4 *>(desc + 8 * (ob + j));
            d = hamming256_csa(a0, a1, __ldg(pb), __ldg(pb + 1));
        }
        const bool hit = d <= thr;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) out[(size_t)r * out_stride + n + __popc(m & ((1u << lane) - 1u))] = (d << 16) | (unsigned)j;
        n += __popc(m);
    }
    if (lane == 0) out_count[r] = n;
}

int run_row_scan(sg_ctx *ctx, const sg_db *db, const int *d_rows, int n_rows, unsigned thr, int out_stride, uint32_t *d_out,
                 uint32_t *d_count) {
    if (n_rows <= 0) return SG_OK;
    row_scan_kernel<<<(n_rows + 3) / 4, 128, 0, ctx->stream>>>(db->d_desc, db->d_offsets, d_rows, n_rows, thr, out_stride, d_out, d_count);
    SG_LAUNCH_CHECK(ctx);
    return SG_OK;
}

int run_hamming(sg_ctx *ctx, const uint32_t *d_a, const uint32_t *d_b, int n, uint32_t *d_out) {
    hamming_pairs_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_a, d_b, n, d_out);
    SG_LAUNCH_CHECK(ctx);
    return SG_OK;
}

int run_popc_bench(sg_ctx *ctx, double *popc_per_s, float *ms_out) {
    unsigned *d_out = nullptr;
    SG_CUDA(ctx, cudaMalloc(&d_out, 4));
    const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        SG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        popc_bench_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 12345u + rep);
        SG_LAUNCH_CHECK(ctx);
        SG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        SG_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        SG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaFree(d_out);
    *popc_per_s = (double)blocks * threads * iters * 8.0 / (best * 1e-3);
    *ms_out = best;
    return SG_OK;
}

}  // namespace sg

extern "C" void sg_angle_bin_order_depth(const uint32_t *sizes30, int depth_limit, uint32_t *order30) {
    sg::Sort30 s;
    s.run(sizes30, depth_limit);
    for (int i = 0; i < 30; ++i) order30[i] = s.v[i];
}
extern "C" void sg_angle_bin_order(const uint32_t *sizes30, uint32_t *order30) {
    sg::Sort30 s;
    s.run(sizes30, 8);
    for (int i = 0; i < 30; ++i) order30[i] = s.v[i];
}
extern "C" int sg_angle_bin(float delta_angle) { return sg::angle_bin(delta_angle); }
