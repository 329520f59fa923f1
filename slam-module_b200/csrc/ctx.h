// Internal context of libslamgpu.so (not part of the ABI; see include/slamgpu.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/slamgpu.h"

namespace sg {

constexpr int PATCH_RADIUS = 19;   // StaticSettings::ORB_PATCH_RADIUS (static_settings.hpp:14)
constexpr int HALF_PATCH = 15;     // ORB_FAST_PATCH_HALF_SIZE (static_settings.hpp:16)
constexpr int CELL = 64;           // FAST cell size of the upstream detector
constexpr int FAST_BORDER = 3;     // cv::FAST never evaluates the outer 3 px of (a sub-)image
constexpr int EVAL_ORIGIN = PATCH_RADIUS + FAST_BORDER;  // first evaluated pixel: 22

// One entry of the horizontal / vertical linear-resize table (cv::resize INTER_LINEAR, 11-bit
// fixed-point coefficients), precomputed on the host exactly as OpenCV does (tables.cpp).
struct ResizeTap {
    int32_t s0, s1;   // source indices (s1 already clipped)
    int16_t a0, a1;   // coefficients, a0 + a1 == 2048
};

struct Level {
    int w = 0, h = 0, pitch = 0;
    size_t frame_stride = 0;        // bytes between consecutive frames of this level
    float scale = 1.f;
    int budget = 0;
    uint8_t *pyr = nullptr;         // [max_frames][h][pitch]; level 0 may alias an external device image
    uint8_t *blur = nullptr;
    ResizeTap *xtab = nullptr;      // [w]   (levels >= 1)
    ResizeTap *ytab = nullptr;      // [h]
    uint4 *xtile = nullptr;         // fast resize kernel: per (tile column, window column pair) {coef a, coef b, byte offset of the 8-byte source window, PRMT selector}
    uint4 *ytile = nullptr;         // ... per (tile row, window row) {offset of source row 0, of source row 1, b0 << 12, b1 << 12}
    bool area2x = false;            // cv::resize switches to INTER_AREA for an exact 2x decimation
    bool fast_resize = false;       // adjacent taps at most 2 source pixels apart: the IDP.2A kernel applies
    int tma_src_w = 0, tma_src_h = 0;    // TMA box of the source tile of the fast resize kernel
    // TMA descriptors (3-D: x, y, frame).  map_src: box over the level BELOW (source of this level's
    // resize; for level 0: box over level 0 itself for the blur-only kernel).  map_fast: 80x70 box over
    // this level for the FAST cells.  Maps over level 0 are re-encoded when the input pointer changes.
    CUtensorMap map_src{}, map_fast{};
    CUtensorMap map_mom{}, map_blur{};   // describe kernel: 48x31 box over this level's plane, 64x37 over its blurred plane
    int src_tile_w = 0, src_tile_h = 0;  // smem extent of the source tile of the resize kernel
    // detection geometry
    int area_w = 0, area_h = 0;     // working area (image minus 19-px border)
    int cells_x = 0, cells_y = 0;   // 64-px cells over the evaluated interior [22, w-22) x [22, h-22)
    int cand_cap = 0;               // candidate capacity per frame
    int node_cap = 0;               // quadtree node capacity == keypoint capacity per frame
    int init_nx = 1, init_ny = 1;   // initial quadtree nodes
    size_t cand_off = 0;            // offset of this level inside a frame's candidate block
    int kp_off = 0;                 // offset of this level inside a frame's detected-keypoint block
};

// Geometry the kernels read (passed by value / __grid_constant__).
struct LevelDev {
    int w, h, pitch;
    unsigned long long frame_stride;
    const uint8_t *pyr;
    const uint8_t *blur;
    float scale;
    int budget;
    int area_w, area_h, cells_x, cells_y;
    int cand_cap, node_cap, init_nx, init_ny;
    unsigned long long cand_off;
    int kp_off;
};
struct GeomDev {
    int levels;
    int det_cap;          // detected keypoints per frame (sum of node_cap)
    int out_cap;          // output keypoints per frame (max_tracks + det_cap)
    int max_tracks;
    unsigned long long cand_per_frame;
    int ini_thr, min_thr;
    int frame0;           // first frame of this launch (pipelined sg_extract works on slices of the batch)
    LevelDev lv[SG_MAX_LEVELS];
};

}  // namespace sg

struct sg_db {
    sg_ctx *ctx = nullptr;
    uint32_t *d_desc = nullptr;
    float *d_angle = nullptr;
    long long *d_offsets = nullptr;
    std::vector<long long> offsets;
    int n_sets = 0;
    int max_set = 0;
    bool owns_data = true;   // false: d_desc / d_angle belong to the caller (sg_db_wrap_device)
};

struct sg_ctx {
    int device = 0;
    sg_params p{};
    cudaStream_t stream = nullptr;      // stream the stage launchers use (swapped per chunk by the pipelined sg_extract)
    static constexpr int N_CMP = 8;
    cudaStream_t main_stream = nullptr, s_in = nullptr, s_out = nullptr, s_cmp[N_CMP] = {};
    int pipe_streams = 4;               // compute streams the chunks rotate over (up to N_CMP; more than 4 measured no gain)
    int overlap_parts = 4;              // sg_extract_device: independent slices of the batch ...
    int overlap_streams = 4;            // ... rotating over this many compute streams
    std::vector<cudaEvent_t> pipe_ev;   // [2 * chunks]: H2D done, compute done
    static constexpr int N_TICKETS = 8;  // sg_extract_submit: completion events of the batches in flight
    cudaEvent_t ticket_ev[N_TICKETS] = {};
    struct Ticket { bool busy = false; int base = 0, n = 0; } ticket[N_TICKETS];
    int next_ticket = 0;
    int stream_chunk = 128;              // chunk of sg_extract_submit (no fill / drain ramp)
    cudaMemPool_t pool = nullptr;        // stream-ordered pool of this context (scratch, descriptor databases)
    bool bow_attr_set = false;           // dynamic shared memory limit of bow_score_kernel raised on this device
    unsigned pipe_rr = 0;                // round-robin position over the compute streams
    int pipe_chunk = 32;                // frames per pipeline chunk of sg_extract
    int frame0 = 0;                     // first frame the stage launchers work on
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr;
    bool in_pipeline = false;           // stage events are not recorded inside the pipelined sg_extract
    std::string err;
    unsigned long long launches = 0;
    // optional per-stage CUDA-event timing (sg_set_profiling): pyramid, fast, distribute, describe,
    // match top-K, match resolve
    bool profiling = false;
    static constexpr int PROF_SLOTS = 64;           // calls remembered (ring)
    cudaEvent_t ev_stage[PROF_SLOTS][8] = {};
    unsigned char stage_mark[PROF_SLOTS][8] = {};
    long prof_call = -1;                             // index of the current call since profiling was switched on
    int sm_count = 148;
    int describe_ctas_per_sm = 0;      // occupancy of the persistent describe kernel (queried on first use)
    bool dist_carveout_set = false;    // the same for distribute_kernel
    bool fast_carveout_set = false;    // fast_cells_kernel's shared-memory carve-out preference (per device: set once per context)

    std::vector<sg_db *> dbs;           // descriptor databases created on this context and not yet destroyed
    std::vector<sg::Level> lv;
    sg::GeomDev geom{};
    const uint8_t *level0 = nullptr;   // current level-0 planes (own buffer or caller's device images)
    int level0_pitch = 0;
    size_t level0_stride = 0;
    int frames_ready = 0;              // frames in the current pyramid
    int level0_frames = 0;             // frames the level-0 TMA descriptors cover
    bool detected = false;

    // detection scratch
    unsigned long long *d_cand = nullptr;  // [max_frames][cand_per_frame]  resp<<32 | order key
    uint32_t *d_cand_node = nullptr;       // node id of each candidate during distribution
    int *d_cand_count = nullptr;           // [max_frames][levels]
    int *d_kp_xy = nullptr;                // [max_frames][det_cap]  x | y<<16 (level coords)
    int *d_kp_resp = nullptr;
    int *d_kp_count = nullptr;             // [max_frames][levels]
    int *d_err = nullptr;                  // device-side overflow flags: [0] synchronous calls, [1 + t] batch of ticket t
    int *h_err = nullptr;                  // pinned mirror read back behind each streamed batch
    int err_slot = 0;                      // which flag the stage launchers hand to the kernels
    int4 *d_cell_table = nullptr;          // FAST cells: level, cell row / column, origin, extent (fast_cell_table)

    // tracker points (host-filtered, orb_extractor.cpp:89-104)
    int *d_trk_xy = nullptr;               // [max_frames][max_tracks]  x | y<<16 at track_level
    float *d_trk_pt = nullptr;             // [max_frames][max_tracks][2] original full-res point
    int *d_trk_id = nullptr;
    int *d_trk_count = nullptr;            // [max_frames]
    bool have_tracks = false;

    // extraction output (SoA, out_cap per frame)
    float *d_x = nullptr, *d_y = nullptr, *d_angle = nullptr;
    int *d_octave = nullptr, *d_track_id = nullptr, *d_lvl_x = nullptr, *d_lvl_y = nullptr;
    uint32_t *d_desc = nullptr;
    int *d_count = nullptr;                // [max_frames]

    // staging
    uint32_t *d_pack = nullptr, *h_pack = nullptr;   // packed outputs of the small-batch latency path (extract_lean)
    uint8_t *h_stage = nullptr;            // pinned, one batch of images
    size_t h_stage_bytes = 0;
    void *d_flush = nullptr;
    size_t flush_bytes = 0;

    // matcher scratch (grown on demand)
    uint32_t *d_topk = nullptr;            // [rows][4] keys
    uint32_t *d_nseen = nullptr;           // [rows]
    size_t topk_rows = 0, topk_words = 0;   // capacity in rows (counts) and in 4-byte keys
    int *d_pairs = nullptr;
    size_t pairs_cap = 0;
    int *d_matches = nullptr;
    size_t matches_cap = 0;
    uint32_t *d_nmatch = nullptr;
    size_t nmatch_cap = 0;
    unsigned long long *d_rescans = nullptr;
    unsigned long long rescans = 0;
    void *d_tmp = nullptr;                 // small uploads for the single-pair entry points
    size_t tmp_bytes = 0;
    void *d_dbtmp = nullptr;               // transient descriptor database of the single-call matchers
    size_t dbtmp_bytes = 0;
};

namespace sg {

// Error plumbing ---------------------------------------------------------------------------------
int fail(sg_ctx *ctx, int code, const char *fmt, ...);
#define SG_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return sg::fail((ctx), SG_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                  \
                            cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)
#define SG_LAUNCH_CHECK(ctx)                                                                      \
    do {                                                                                          \
        ++(ctx)->launches;                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess)                                                                    \
            return sg::fail((ctx), SG_ERR_CUDA, "kernel launch failed: %s (%s:%d)",              \
                            cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)

// Host-side tables (tables.cpp) -----------------------------------------------------------------
void make_geometry(const sg_params &p, std::vector<Level> &lv);
void make_resize_taps(int src, int dst, std::vector<ResizeTap> &taps, bool horizontal);
bool is_area2x(int sw, int sh, int dw, int dh);

// Stage launchers (one per .cu) -----------------------------------------------------------------
int launch_pyramid(sg_ctx *ctx, int n_frames);
int launch_detect(sg_ctx *ctx, int n_frames);
int launch_describe(sg_ctx *ctx, int n_frames);
void describe_box_dims(int *mom_w, int *mom_h, int *blur_w, int *blur_h);
int grow(sg_ctx *ctx, void **ptr, size_t *cap, size_t need, size_t elem);
// Device scratch carved from the context's grow-only buffer (no cudaMalloc / cudaFree per call).
struct Scratch {
    sg_ctx *ctx;
    size_t need = 0, at = 0;
    explicit Scratch(sg_ctx *c) : ctx(c) {}
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    void want(size_t bytes) { need += pad(std::max<size_t>(bytes, 1)); }
    int commit() { return grow(ctx, &ctx->d_tmp, &ctx->tmp_bytes, need, 1); }
    template <class T> T *take(size_t n) {
        T *p = reinterpret_cast<T *>(static_cast<uint8_t *>(ctx->d_tmp) + at);
        at += pad(std::max<size_t>(n * sizeof(T), 1));
        return p;
    }
    template <class T> int put(T **d, const T *h, size_t n) {
        *d = take<T>(n);
        if (n) SG_CUDA(ctx, cudaMemcpyAsync(*d, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        return SG_OK;
    }
};
// TMA descriptor of an 8-bit plane stack {w, h, frames} with a fixed box (tma.cpp).
int encode_plane_map(sg_ctx *ctx, CUtensorMap *out, const uint8_t *base, int w, int h, int pitch, size_t frame_stride,
                     int frames, int box_w, int box_h, bool swizzle64 = false);
int encode_level0_maps(sg_ctx *ctx);
// Record stage event i on the context's stream when profiling is on.
inline void mark(sg_ctx *ctx, int i, bool first = false) {
    if (!ctx->profiling || ctx->in_pipeline) return;
    if (first) {   // a new call: take the next ring slot
        ++ctx->prof_call;
        for (auto &m : ctx->stage_mark[ctx->prof_call % sg_ctx::PROF_SLOTS]) m = 0;
    }
    if (ctx->prof_call < 0) return;
    const int slot = (int)(ctx->prof_call % sg_ctx::PROF_SLOTS);
    cudaEventRecord(ctx->ev_stage[slot][i], ctx->stream);
    ctx->stage_mark[slot][i] = 1;
}
enum { EV_PYR0 = 0, EV_PYR1, EV_FAST1, EV_DIST1, EV_DESC1, EV_MATCH0, EV_TOPK1, EV_RESOLVE1 };

}  // namespace sg
