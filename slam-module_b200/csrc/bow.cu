// Bag-of-words row of the hot path: what sits between extraction and loop-closure matching.
//
//   BowIndex::transform      bow_index.cpp:59-93    DBoW2 TemplatedVocabulary::transform(features, bowVector,
//                                                   featureVector, levelsUp = 4): tree descent per descriptor,
//                                                   BowVector = sum of word weights in feature order, L1 normalised
//   BowIndex::add / remove   bow_index.cpp:44-57    keyframe -> inverted lists
//   BowIndex::getBowSimilar  bow_index.cpp:95-176   words in common, DBoW2 L1 score, best-score window
//
// DBoW2 is an external dependency of the reference (absent from its tree); the arithmetic restated here is the
// published one: WordValue = double, TF-IDF weighting, L1 norm and L1Scoring::score
// (sum over the common words in ascending word order of |v - w| - |v| - |w|, then -sum / 2).
//
// GPU formulation (exact, including the order of every double addition):
//   bow_transform_kernel   one warp per feature, lanes over the children of the current node.
//   bow_vector_kernel      one CTA per keyframe: bitonic sort of (word, feature index), one thread per word adds the
//                          weights of its run in feature order, one thread accumulates the L1 norm in word order,
//                          all threads divide.
//   bow_score_kernel       no inverted lists: every stored keyframe is scored against the query by one warp
//                          (lanes binary-search the query's words; matched terms are added in lane = word order),
//                          which yields words-in-common and the score for all keyframes in one launch.  The
//                          selection rules of getBowSimilar then run on the host over these two arrays, in the
//                          reference's std::map order and with the same std::sort.
#include <algorithm>
#include <map>
#include <utility>
#include <vector>
#include "ctx.h"

namespace sg {

constexpr int BOW_WARPS = 8;
constexpr int BOWV_MAX = 4096;        // features per keyframe handled by bow_vector_kernel
constexpr int BOWV_THREADS = 1024;

__device__ __forceinline__ unsigned hamming8(const uint32_t (&a)[8], const uint32_t *b) {
    const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(b)), b1 = __ldg(reinterpret_cast<const uint4 *>(b) + 1);
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w)
           + __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// ---- BoW transform: DBoW2 vocabulary-tree descent (bow_index.cpp:59-93 -> TemplatedVocabulary::transform) ---------
// One warp per feature: at every level the lanes take the children of the current node (32 at a time), each computes
// one Hamming distance, the warp minimum of (distance << 16 | child position) picks the nearest child with DBoW2's
// tie rule (strict '<': the first child wins).  The word is the leaf's; the feature-vector node is the one reached at
// level L - levelsUp.
struct VocabDev {
    const int *child_off, *child_ids, *node_word;
    const uint32_t *node_desc;
    const double *node_weight;
    int levels;
};

__global__ void __launch_bounds__(BOW_WARPS * 32)
bow_transform_kernel(const VocabDev v, const uint32_t *desc, int n, int levels_up, int *out_word, double *out_weight, int *out_node) {
    const int lane = threadIdx.x & 31, f = blockIdx.x * BOW_WARPS + (threadIdx.x >> 5);
    if (f >= n) return;
    uint32_t d[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) d[w] = __ldg(desc + 8 * (size_t)f + w);
    const int nid_level = v.levels - levels_up;
    int cur = 0, level = 0, nid = 0;
    while (true) {
        const int b = __ldg(v.child_off + cur), e = __ldg(v.child_off + cur + 1);
        if (e <= b) break;
        ++level;
        unsigned best = 0xffffffffu;
        for (int c0 = b; c0 < e; c0 += 32) {          // position inside the child list orders the ties
            const int c = c0 + lane;
            unsigned key = 0xffffffffu;
            if (c < e) key = (hamming8(d, v.node_desc + 8 * (size_t)__ldg(v.child_ids + c)) << 16) | (unsigned)min(c - b, 0xffff);
            best = min(best, __reduce_min_sync(0xffffffffu, key));
        }
        cur = __ldg(v.child_ids + b + (int)(best & 0xffffu));
        if (level == nid_level) nid = cur;
    }
    if (lane == 0) {
        out_word[f] = v.node_word[cur];
        out_weight[f] = v.node_weight[cur];
        out_node[f] = nid_level <= 0 ? 0 : nid;
    }
}

// ---- BowVector of one keyframe (TemplatedVocabulary::transform, TF-IDF branch + BowVector::normalize(L1)) ---------
// v.addWeight(word, w) for every feature with w > 0, in feature order: the value of a word is the left-to-right sum
// of its features' weights; normalize: norm = sum of |value| in ascending word order, value /= norm when norm > 0.
__global__ void __launch_bounds__(BOWV_THREADS)
bow_vector_kernel(const int *word, const double *weight, const long long *offsets, uint32_t *vec_word, double *vec_value,
                  int *n_words) {
    __shared__ unsigned long long key[BOWV_MAX];
    __shared__ int warp_tot[BOWV_THREADS / 32];
    __shared__ double s_norm;
    const int t = threadIdx.x;
    // keyframe blockIdx.x owns features [offsets[k], offsets[k + 1]); its BowVector goes to the same offset
    const long long o = offsets[blockIdx.x];
    const int n = (int)(offsets[blockIdx.x + 1] - o);
    word += o; weight += o; vec_word += o; vec_value += o; n_words += blockIdx.x;
    for (int i = t; i < BOWV_MAX; i += BOWV_THREADS)
        key[i] = (i < n && weight[i] > 0.0) ? ((unsigned long long)(unsigned)word[i] << 32 | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= BOWV_MAX; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < BOWV_MAX; i += BOWV_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = key[i], b = key[p];
                    if ((a > b) == ((i & k) == 0)) { key[i] = b; key[p] = a; }
                }
            }
            __syncthreads();
        }
    // heads of the word runs, 4 consecutive positions per thread, exclusive scan over the CTA
    constexpr int PER = BOWV_MAX / BOWV_THREADS;
    int head[PER], cnt = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int p = PER * t + u;
        const unsigned long long k = key[p];
        head[u] = k != ~0ull && (p == 0 || (unsigned)(key[p - 1] >> 32) != (unsigned)(k >> 32));
        cnt += head[u];
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((t & 31) >= d) incl += o;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    int base = incl - cnt;
    for (int w = 0; w < (t >> 5); ++w) base += warp_tot[w];
    int total = 0;
    for (int w = 0; w < BOWV_THREADS / 32; ++w) total += warp_tot[w];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        if (!head[u]) continue;
        int p = PER * t + u;
        const unsigned wd = (unsigned)(key[p] >> 32);
        double sum = weight[(unsigned)key[p]];
        for (++p; p < BOWV_MAX && (unsigned)(key[p] >> 32) == wd && key[p] != ~0ull; ++p) sum += weight[(unsigned)key[p]];
        vec_word[base] = wd;
        vec_value[base] = sum;
        ++base;
    }
    __syncthreads();
    if (t == 0) {
        double norm = 0.0;
        for (int i = 0; i < total; ++i) norm += fabs(vec_value[i]);
        s_norm = norm;
        *n_words = total;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0)
        for (int i = t; i < total; i += BOWV_THREADS) vec_value[i] /= norm;
}

// ---- words in common + L1 score of the query against every stored keyframe (bow_index.cpp:104-150) -----------------
__global__ void __launch_bounds__(BOW_WARPS * 32)
bow_score_kernel(const uint32_t *db_word, const double *db_value, const int *db_len, int stride, int n_slots,
                 const uint32_t *q_word, const double *q_value, int nq, uint32_t *common, float *score) {
    extern __shared__ __align__(16) unsigned char bow_smem[];
    double *sv = reinterpret_cast<double *>(bow_smem);
    uint32_t *sw = reinterpret_cast<uint32_t *>(sv + nq);
    for (int i = threadIdx.x; i < nq; i += blockDim.x) { sv[i] = q_value[i]; sw[i] = q_word[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, s = blockIdx.x * BOW_WARPS + (threadIdx.x >> 5);
    if (s >= n_slots) return;
    const int len = db_len[s];
    const uint32_t *w = db_word + (size_t)s * stride;
    const double *v = db_value + (size_t)s * stride;
    double acc = 0.0;
    unsigned n_common = 0;
    for (int i0 = 0; i0 < len; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false;
        double term = 0.0;
        if (i < len) {
            const uint32_t wd = __ldg(w + i);
            int lo = 0, hi = nq;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sw[mid] < wd) lo = mid + 1; else hi = mid;
            }
            if (lo < nq && sw[lo] == wd) {
                hit = true;
                const double vi = sv[lo], wi = __ldg(v + i);      // score(query, stored): vi from the query
                term = fabs(vi - wi) - fabs(vi) - fabs(wi);
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        n_common += __popc(m);
        while (m) {                                                // ascending word order, one addition per common word
            const int l = __ffs(m) - 1;
            acc += __shfl_sync(0xffffffffu, term, l);
            m &= m - 1;
        }
    }
    if (lane == 0) {
        common[s] = n_common;
        score[s] = (float)(-acc / 2.0);
    }
}

}  // namespace sg

using namespace sg;

struct sg_vocab {
    sg_ctx *ctx = nullptr;
    int *child_off = nullptr, *child_ids = nullptr, *node_word = nullptr;
    uint32_t *node_desc = nullptr;
    double *node_weight = nullptr;
    int n_nodes = 0, levels = 0;
};

extern "C" void sg_vocab_destroy(sg_vocab *v) {
    if (!v) return;
    if (v->ctx) cudaSetDevice(v->ctx->device);
    cudaFree(v->child_off); cudaFree(v->child_ids); cudaFree(v->node_word); cudaFree(v->node_desc); cudaFree(v->node_weight);
    delete v;
}

extern "C" int sg_vocab_create(sg_ctx *ctx, const int32_t *h_child_off, const int32_t *h_child_ids, const uint32_t *h_node_desc,
                               const double *h_node_weight, const int32_t *h_node_word, int n_nodes, int levels, sg_vocab **out) {
    cudaSetDevice(ctx->device);
    if (!h_child_off || !h_node_desc || !h_node_weight || !h_node_word || n_nodes < 1 || levels < 0 || !out)
        return fail(ctx, SG_ERR_INVALID, "null / empty vocabulary");
    const int n_children = h_child_off[n_nodes];
    if (h_child_off[0] != 0 || n_children < 0 || (n_children && !h_child_ids)) return fail(ctx, SG_ERR_INVALID, "bad child offsets");
    for (int i = 0; i < n_nodes; ++i) {
        if (h_child_off[i + 1] < h_child_off[i]) return fail(ctx, SG_ERR_INVALID, "child offsets must be non-decreasing");
        if (h_child_off[i + 1] - h_child_off[i] > 65535) return fail(ctx, SG_ERR_INVALID, "more than 65535 children under one node");
    }
    for (int c = 0; c < n_children; ++c)   // a child id must be a later node: the descent then terminates on any input
        if (h_child_ids[c] <= 0 || h_child_ids[c] >= n_nodes) return fail(ctx, SG_ERR_INVALID, "child id %d outside the tree", h_child_ids[c]);
    for (int i = 0; i < n_nodes; ++i)
        for (int c = h_child_off[i]; c < h_child_off[i + 1]; ++c)
            if (h_child_ids[c] <= i) return fail(ctx, SG_ERR_INVALID, "node %d lists child %d: children must have larger ids than their parent", i, h_child_ids[c]);
    sg_vocab *v = new sg_vocab();
    v->ctx = ctx; v->n_nodes = n_nodes; v->levels = levels;
    auto put = [&](auto **d, const auto *h, size_t n) -> bool {
        if (cudaMalloc((void **)d, std::max<size_t>(n, 1) * sizeof(**d)) != cudaSuccess) return false;
        return n == 0 || cudaMemcpy(*d, h, n * sizeof(**d), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!put(&v->child_off, h_child_off, (size_t)n_nodes + 1) || !put(&v->child_ids, h_child_ids, (size_t)n_children)
        || !put(&v->node_desc, h_node_desc, 8 * (size_t)n_nodes) || !put(&v->node_weight, h_node_weight, (size_t)n_nodes)
        || !put(&v->node_word, h_node_word, (size_t)n_nodes)) {
        sg_vocab_destroy(v);
        return fail(ctx, SG_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = v;
    return SG_OK;
}

static int bow_launch(sg_ctx *ctx, const sg_vocab *v, const uint32_t *d_desc, int n, int levels_up, int *d_word, double *d_weight, int *d_node) {
    VocabDev dv{v->child_off, v->child_ids, v->node_word, v->node_desc, v->node_weight, v->levels};
    bow_transform_kernel<<<(n + BOW_WARPS - 1) / BOW_WARPS, BOW_WARPS * 32, 0, ctx->stream>>>(dv, d_desc, n, levels_up, d_word, d_weight, d_node);
    SG_LAUNCH_CHECK(ctx);
    return SG_OK;
}

extern "C" int sg_bow_transform(sg_ctx *ctx, const sg_vocab *vocab, const uint32_t *h_desc, int n, int levels_up, int32_t *h_word,
                                double *h_weight, int32_t *h_node) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return SG_OK;
    if (!vocab || !h_desc || !h_word || !h_weight || !h_node) return fail(ctx, SG_ERR_INVALID, "null argument");
    Scratch sc(ctx);
    sc.want(32 * (size_t)n); sc.want(4 * (size_t)n); sc.want(8 * (size_t)n); sc.want(4 * (size_t)n);
    if (int r = sc.commit()) return r;
    uint32_t *d_desc;
    if (int r = sc.put(&d_desc, h_desc, 8 * (size_t)n)) return r;
    int *d_word = sc.take<int>(n);
    double *d_weight = sc.take<double>(n);
    int *d_node = sc.take<int>(n);
    if (int r = bow_launch(ctx, vocab, d_desc, n, levels_up, d_word, d_weight, d_node)) return r;
    SG_CUDA(ctx, cudaMemcpyAsync(h_word, d_word, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(h_weight, d_weight, 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(h_node, d_node, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

// Device-resident form: descriptors already on the device (e.g. sg_extract_device_views), results stay there.
extern "C" int sg_bow_transform_device(sg_ctx *ctx, const sg_vocab *vocab, const uint32_t *d_desc, int n, int levels_up,
                                       int32_t *d_word, double *d_weight, int32_t *d_node) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return SG_OK;
    if (!vocab || !d_desc || !d_word || !d_weight || !d_node) return fail(ctx, SG_ERR_INVALID, "null argument");
    return bow_launch(ctx, vocab, d_desc, n, levels_up, d_word, d_weight, d_node);
}

// ---- BowVector, keyframe database, similarity query ----------------------------------------------------------
extern "C" int sg_bow_vector_batch(sg_ctx *ctx, const int32_t *h_word, const double *h_weight, const int64_t *h_offsets,
                                   int n_keyframes, uint32_t *h_vec_word, double *h_vec_value, int32_t *n_words) {
    cudaSetDevice(ctx->device);
    if (n_keyframes <= 0) return SG_OK;
    if (!h_offsets || !n_words) return fail(ctx, SG_ERR_INVALID, "null argument");
    if (h_offsets[0] != 0) return fail(ctx, SG_ERR_INVALID, "offsets[0] must be 0");
    for (int k = 0; k < n_keyframes; ++k) {
        const long long n = h_offsets[k + 1] - h_offsets[k];
        if (n < 0 || n > BOWV_MAX) return fail(ctx, SG_ERR_INVALID, "keyframe %d has %lld features (supported: 0..%d)", k, n, BOWV_MAX);
        n_words[k] = 0;
    }
    const size_t total = (size_t)h_offsets[n_keyframes];
    if (total == 0) return SG_OK;
    if (!h_word || !h_weight || !h_vec_word || !h_vec_value) return fail(ctx, SG_ERR_INVALID, "null argument");
    Scratch sc(ctx);
    sc.want(4 * total); sc.want(8 * total); sc.want(4 * total); sc.want(8 * total);
    sc.want(8 * ((size_t)n_keyframes + 1)); sc.want(4 * (size_t)n_keyframes);
    if (int r = sc.commit()) return r;
    int *d_word;
    double *d_weight;
    long long *d_off;
    if (int r = sc.put(&d_word, (const int *)h_word, total)) return r;
    if (int r = sc.put(&d_weight, h_weight, total)) return r;
    uint32_t *d_vw = sc.take<uint32_t>(total);
    double *d_vv = sc.take<double>(total);
    if (int r = sc.put(&d_off, (const long long *)h_offsets, (size_t)n_keyframes + 1)) return r;
    int *d_n = sc.take<int>(n_keyframes);
    bow_vector_kernel<<<n_keyframes, BOWV_THREADS, 0, ctx->stream>>>(d_word, d_weight, d_off, d_vw, d_vv, d_n);
    SG_LAUNCH_CHECK(ctx);
    SG_CUDA(ctx, cudaMemcpyAsync(n_words, d_n, 4 * (size_t)n_keyframes, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(h_vec_word, d_vw, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(h_vec_value, d_vv, 8 * total, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

extern "C" int sg_bow_vector(sg_ctx *ctx, const int32_t *h_word, const double *h_weight, int n, uint32_t *h_vec_word,
                             double *h_vec_value, int *n_words) {
    if (!n_words) return fail(ctx, SG_ERR_INVALID, "null argument");
    *n_words = 0;
    if (n <= 0) return SG_OK;
    if (n > BOWV_MAX) return fail(ctx, SG_ERR_INVALID, "%d features (supported: up to %d per keyframe)", n, BOWV_MAX);
    const int64_t offsets[2] = {0, n};
    return sg_bow_vector_batch(ctx, h_word, h_weight, offsets, 1, h_vec_word, h_vec_value, n_words);
}

struct sg_bowdb {
    sg_ctx *ctx = nullptr;
    int max_keyframes = 0, stride = 0, n_slots = 0;          // n_slots: slots ever used (dead ones have length 0)
    uint32_t *d_word = nullptr, *d_common = nullptr;
    double *d_value = nullptr;
    int *d_len = nullptr;
    float *d_score = nullptr;
    std::map<std::pair<int, int>, int> slot_of;              // (map id, keyframe id) -> slot; the reference's MapKf order
    std::vector<int> free_slots;
    std::vector<uint32_t> h_common;
    std::vector<float> h_score;
};

extern "C" void sg_bowdb_destroy(sg_bowdb *db) {
    if (!db) return;
    if (db->ctx) cudaSetDevice(db->ctx->device);
    cudaFree(db->d_word); cudaFree(db->d_value); cudaFree(db->d_len); cudaFree(db->d_common); cudaFree(db->d_score);
    delete db;
}

extern "C" int sg_bowdb_create(sg_ctx *ctx, int max_keyframes, int max_words_per_keyframe, sg_bowdb **out) {
    cudaSetDevice(ctx->device);
    if (!out || max_keyframes < 1 || max_words_per_keyframe < 1) return fail(ctx, SG_ERR_INVALID, "bad BoW database size");
    sg_bowdb *db = new sg_bowdb();
    db->ctx = ctx; db->max_keyframes = max_keyframes; db->stride = (max_words_per_keyframe + 31) & ~31;
    const size_t cells = (size_t)max_keyframes * db->stride;
    if (cudaMalloc((void **)&db->d_word, 4 * cells) != cudaSuccess || cudaMalloc((void **)&db->d_value, 8 * cells) != cudaSuccess
        || cudaMalloc((void **)&db->d_len, 4 * (size_t)max_keyframes) != cudaSuccess
        || cudaMalloc((void **)&db->d_common, 4 * (size_t)max_keyframes) != cudaSuccess
        || cudaMalloc((void **)&db->d_score, 4 * (size_t)max_keyframes) != cudaSuccess
        || cudaMemset(db->d_len, 0, 4 * (size_t)max_keyframes) != cudaSuccess) {
        sg_bowdb_destroy(db);
        return fail(ctx, SG_ERR_CUDA, "BoW database allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    db->h_common.resize(max_keyframes);
    db->h_score.resize(max_keyframes);
    *out = db;
    return SG_OK;
}

extern "C" int sg_bowdb_size(const sg_bowdb *db) { return db ? (int)db->slot_of.size() : 0; }

// BowIndex::add (bow_index.cpp:44-48): the keyframe's BowVector becomes searchable.
extern "C" int sg_bowdb_add(sg_ctx *ctx, sg_bowdb *db, int map_id, int kf_id, const uint32_t *h_vec_word,
                            const double *h_vec_value, int n_words) {
    cudaSetDevice(ctx->device);
    if (!db || n_words < 0 || (n_words && (!h_vec_word || !h_vec_value))) return fail(ctx, SG_ERR_INVALID, "null argument");
    if (n_words > db->stride) return fail(ctx, SG_ERR_INVALID, "%d words, the database holds %d per keyframe", n_words, db->stride);
    for (int i = 1; i < n_words; ++i)
        if (h_vec_word[i] <= h_vec_word[i - 1]) return fail(ctx, SG_ERR_INVALID, "BowVector words must be strictly ascending");
    const auto key = std::make_pair(map_id, kf_id);
    if (db->slot_of.count(key)) return fail(ctx, SG_ERR_INVALID, "keyframe (%d, %d) is already in the database", map_id, kf_id);
    int slot;
    if (!db->free_slots.empty()) { slot = db->free_slots.back(); db->free_slots.pop_back(); }
    else if (db->n_slots < db->max_keyframes) slot = db->n_slots++;
    else return fail(ctx, SG_ERR_INVALID, "BoW database is full (%d keyframes)", db->max_keyframes);
    const size_t at = (size_t)slot * db->stride;
    if (n_words) {
        SG_CUDA(ctx, cudaMemcpyAsync(db->d_word + at, h_vec_word, 4 * (size_t)n_words, cudaMemcpyHostToDevice, ctx->stream));
        SG_CUDA(ctx, cudaMemcpyAsync(db->d_value + at, h_vec_value, 8 * (size_t)n_words, cudaMemcpyHostToDevice, ctx->stream));
    }
    SG_CUDA(ctx, cudaMemcpyAsync(db->d_len + slot, &n_words, 4, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // the host arrays may be reused by the caller
    db->slot_of[key] = slot;
    return SG_OK;
}

// BowIndex::remove (bow_index.cpp:50-57).  Removing an unknown keyframe is a no-op, as in the reference.
extern "C" int sg_bowdb_remove(sg_ctx *ctx, sg_bowdb *db, int map_id, int kf_id) {
    cudaSetDevice(ctx->device);
    if (!db) return fail(ctx, SG_ERR_INVALID, "null argument");
    const auto it = db->slot_of.find(std::make_pair(map_id, kf_id));
    if (it == db->slot_of.end()) return SG_OK;
    const int zero = 0;
    SG_CUDA(ctx, cudaMemcpyAsync(db->d_len + it->second, &zero, 4, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    db->free_slots.push_back(it->second);
    db->slot_of.erase(it);
    return SG_OK;
}

// BowIndex::getBowSimilar (bow_index.cpp:95-176).
extern "C" int sg_bow_similar(sg_ctx *ctx, sg_bowdb *db, const uint32_t *h_q_word, const double *h_q_value, int nq,
                              int self_map, int self_kf, float min_in_common_ratio, float score_ratio, int32_t *h_map,
                              int32_t *h_kf, float *h_score, int capacity, int *n_out) {
    cudaSetDevice(ctx->device);
    if (!db || !n_out || capacity < 0 || (capacity && (!h_map || !h_kf || !h_score))) return fail(ctx, SG_ERR_INVALID, "null argument");
    *n_out = 0;
    if (nq <= 0 || db->n_slots == 0) return SG_OK;
    if (!h_q_word || !h_q_value) return fail(ctx, SG_ERR_INVALID, "null argument");
    if (nq > BOWV_MAX) return fail(ctx, SG_ERR_INVALID, "%d query words (supported: up to %d)", nq, BOWV_MAX);
    for (int i = 1; i < nq; ++i)
        if (h_q_word[i] <= h_q_word[i - 1]) return fail(ctx, SG_ERR_INVALID, "BowVector words must be strictly ascending");
    Scratch sc(ctx);
    sc.want(4 * (size_t)nq); sc.want(8 * (size_t)nq);
    if (int r = sc.commit()) return r;
    uint32_t *d_qw;
    double *d_qv;
    if (int r = sc.put(&d_qw, h_q_word, (size_t)nq)) return r;
    if (int r = sc.put(&d_qv, h_q_value, (size_t)nq)) return r;
    const size_t smem = 12 * (size_t)nq;
    if (!ctx->bow_attr_set) {       // per device, so per context
        SG_CUDA(ctx, cudaFuncSetAttribute(bow_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * BOWV_MAX));
        ctx->bow_attr_set = true;
    }
    bow_score_kernel<<<(db->n_slots + BOW_WARPS - 1) / BOW_WARPS, BOW_WARPS * 32, smem, ctx->stream>>>(
        db->d_word, db->d_value, db->d_len, db->stride, db->n_slots, d_qw, d_qv, nq, db->d_common, db->d_score);
    SG_LAUNCH_CHECK(ctx);
    SG_CUDA(ctx, cudaMemcpyAsync(db->h_common.data(), db->d_common, 4 * (size_t)db->n_slots, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(db->h_score.data(), db->d_score, 4 * (size_t)db->n_slots, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    // :128-136 the best word count over every keyframe that shares a word (std::map order = (map id, keyframe id))
    struct Similar { int map, kf; float score; };
    unsigned max_in_common = 0;
    bool any = false;
    const auto self = std::make_pair(self_map, self_kf);
    for (const auto &e : db->slot_of) {
        if (e.first == self || db->h_common[e.second] == 0) continue;
        any = true;
        max_in_common = std::max(max_in_common, db->h_common[e.second]);
    }
    if (!any) return SG_OK;
    const unsigned min_in_common = static_cast<unsigned>(min_in_common_ratio * static_cast<float>(max_in_common));   // :139-140
    std::vector<Similar> similar;
    for (const auto &e : db->slot_of) {                                                                               // :142-158
        if (e.first == self || db->h_common[e.second] == 0) continue;
        if (db->h_common[e.second] > min_in_common) similar.push_back(Similar{e.first.first, e.first.second, db->h_score[e.second]});
    }
    if (similar.empty()) return SG_OK;
    std::sort(similar.begin(), similar.end(), [](const Similar &a, const Similar &b) { return a.score > b.score; });   // :163
    const float min_score = similar[0].score * score_ratio;                                                           // :166
    const auto cut = std::find_if(similar.begin(), similar.end(), [&](const Similar &p) { return p.score < min_score; });
    similar.erase(cut, similar.end());
    *n_out = (int)similar.size();
    for (int i = 0; i < std::min(*n_out, capacity); ++i) { h_map[i] = similar[i].map; h_kf[i] = similar[i].kf; h_score[i] = similar[i].score; }
    return SG_OK;
}
