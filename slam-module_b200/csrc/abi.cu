// C ABI of libslamgpu.so (include/slamgpu.h): context life cycle, buffer ownership, host <-> device
// staging and the orchestration of the stage launchers.  No algorithmic work happens here.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include "ctx.h"

namespace sg {

static thread_local std::string g_create_error;

int fail(sg_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

int grow(sg_ctx *ctx, void **ptr, size_t *cap, size_t need, size_t elem) {
    if (need <= *cap) return SG_OK;
    // stream-ordered pool allocation (the pool keeps what is freed): a growing scratch buffer must not cost a
    // device-wide cudaFree / cudaMalloc, which takes milliseconds to seconds on virtualised hosts
    if (*ptr) cudaFreeAsync(*ptr, ctx->main_stream);
    *ptr = nullptr; *cap = 0;
    SG_CUDA(ctx, cudaMallocFromPoolAsync(ptr, need * elem, ctx->pool, ctx->main_stream));
    *cap = need;
    return SG_OK;
}

void pyramid_source_extent(const std::vector<ResizeTap> &xt, const std::vector<ResizeTap> &yt, int w, int h,
                           bool area2x, int sw, int sh, int *tile_w, int *tile_h);
int pyramid_fast_source_rows(const std::vector<ResizeTap> &yt, int h);
void pyramid_tile_tables(const std::vector<ResizeTap> &xt, const std::vector<ResizeTap> &yt, int w, int h, int src_pitch,
                         std::vector<uint4> &xtile, std::vector<uint4> &ytile);
int run_match(sg_ctx *ctx, const sg_db *db, const int *d_pairs, int n_pairs, const sg_match_params &mp,
              int *d_matches, int match_stride, uint32_t *d_n_matches, bool reset_rescans = true);
int match_chunk_pairs(const sg_db *db, bool own_matches);
int run_topk_lists(sg_ctx *ctx, const sg_db *db, const int *d_pairs, int n_pairs, unsigned thr, uint32_t **d_topk,
                   uint32_t **d_nseen, int *row_stride);
int run_row_scan(sg_ctx *ctx, const sg_db *db, const int *d_rows, int n_rows, unsigned thr, int out_stride, uint32_t *d_out,
                 uint32_t *d_count);
int run_hamming(sg_ctx *ctx, const uint32_t *d_a, const uint32_t *d_b, int n, uint32_t *d_out);
int run_popc_bench(sg_ctx *ctx, double *popc_per_s, float *ms_out);
size_t distribute_smem_bytes(int node_cap_max, bool pack);
void fast_cell_table(const GeomDev &g, std::vector<int4> &cells);

template <class T>
static int dev_alloc(sg_ctx *ctx, T **p, size_t n) {
    SG_CUDA(ctx, cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return SG_OK;
}

static int check_device_error(sg_ctx *ctx) {
    int e = 0;
    SG_CUDA(ctx, cudaMemcpyAsync(&e, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (e) {
        cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream);
        return fail(ctx, e, "device-side capacity exceeded (candidate list or quadtree node table)");
    }
    return SG_OK;
}

static int build_context(sg_ctx *ctx) {
    const sg_params &p = ctx->p;
    make_geometry(p, ctx->lv);
    GeomDev &g = ctx->geom;
    memset(&g, 0, sizeof g);
    g.levels = p.levels;
    g.max_tracks = p.max_tracks;
    g.ini_thr = p.ini_fast_thr;
    g.min_thr = p.min_fast_thr;
    size_t cand_off = 0;
    int kp_off = 0, nc_max = 1;
    const size_t F = p.max_frames;
    for (int l = 0; l < p.levels; ++l) {
        Level &L = ctx->lv[l];
        if (L.w < 8 || L.h < 8) return fail(ctx, SG_ERR_INVALID, "pyramid level %d is %dx%d: too small", l, L.w, L.h);
        if (L.area_w > 32000 || L.area_h > 32000) return fail(ctx, SG_ERR_INVALID, "image too large");
        if (L.cells_x > 1023 || L.cells_y > 1023) return fail(ctx, SG_ERR_INVALID, "image too large");
        L.cand_cap = L.cells_x * L.cells_y * (CELL / 2) * (CELL / 2) + 32;   // NMS: <= one pixel per 2x2 block of a cell
        L.cand_off = cand_off; cand_off += (size_t)L.cand_cap;
        L.kp_off = kp_off; kp_off += L.node_cap;
        nc_max = std::max(nc_max, L.node_cap);
        if (int r = dev_alloc(ctx, &L.pyr, F * L.frame_stride + 256)) return r;
        if (int r = dev_alloc(ctx, &L.blur, F * L.frame_stride + 256)) return r;
        if (l > 0) {
            const Level &P = ctx->lv[l - 1];
            std::vector<ResizeTap> xt, yt;
            make_resize_taps(P.w, L.w, xt, true);
            make_resize_taps(P.h, L.h, yt, false);
            L.area2x = is_area2x(P.w, P.h, L.w, L.h);
            if (L.w == P.w && L.h == P.h) {   // cv::resize copies: identity taps
                for (int i = 0; i < L.w; ++i) xt[i] = ResizeTap{i, i, 2048, 0};
                for (int i = 0; i < L.h; ++i) yt[i] = ResizeTap{i, i, 2048, 0};
            }
            pyramid_source_extent(xt, yt, L.w, L.h, L.area2x, P.w, P.h, &L.src_tile_w, &L.src_tile_h);
            L.fast_resize = !L.area2x && P.h < 32768;
            {   // TMA box of the source tile: origin aligned down to 16 bytes (TMA requirement), wide enough
                // for the 8-byte window the kernel reads at the last column's first tap, rounded to 16
                int mw = 0;
                for (int x0 = 0; x0 < L.w; x0 += 64) {
                    const int tw = std::min(64, L.w - x0), xlo = std::max(x0 - 3, 0), xhi = std::min(x0 + tw + 3, L.w);
                    mw = std::max(mw, xt[xhi - 1].s0 - (xt[xlo].s0 & ~15) + 8);
                }
                L.tma_src_w = (mw + 15) & ~15;
                L.tma_src_h = pyramid_fast_source_rows(yt, L.h);
                if (L.tma_src_w > 256 || L.tma_src_h > 255) L.fast_resize = false;
            }
            for (int i = 0; i + 1 < L.w; ++i)
                if (xt[i + 1].s0 - xt[i].s0 > 2 || xt[i + 1].s0 < xt[i].s0) L.fast_resize = false;
            if (L.fast_resize) {
                // Everything the fast resize kernel derives from the taps depends on (level, tile column, window column) or
                // (level, tile row, window row) only -- never on the frame: precomputed here, one 16-byte load per thread there.
                std::vector<uint4> xtile, ytile;
                pyramid_tile_tables(xt, yt, L.w, L.h, L.tma_src_w, xtile, ytile);
                if (int r = dev_alloc(ctx, &L.xtile, xtile.size())) return r;
                if (int r = dev_alloc(ctx, &L.ytile, ytile.size())) return r;
                SG_CUDA(ctx, cudaMemcpy(L.xtile, xtile.data(), xtile.size() * sizeof(uint4), cudaMemcpyHostToDevice));
                SG_CUDA(ctx, cudaMemcpy(L.ytile, ytile.data(), ytile.size() * sizeof(uint4), cudaMemcpyHostToDevice));
            }
            if (int r = dev_alloc(ctx, &L.xtab, xt.size())) return r;
            if (int r = dev_alloc(ctx, &L.ytab, yt.size())) return r;
            SG_CUDA(ctx, cudaMemcpy(L.xtab, xt.data(), xt.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
            SG_CUDA(ctx, cudaMemcpy(L.ytab, yt.data(), yt.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
        }
        LevelDev &D = g.lv[l];
        D.w = L.w; D.h = L.h; D.pitch = L.pitch; D.frame_stride = L.frame_stride;
        D.pyr = L.pyr; D.blur = L.blur; D.scale = L.scale; D.budget = L.budget;
        D.area_w = L.area_w; D.area_h = L.area_h; D.cells_x = L.cells_x; D.cells_y = L.cells_y;
        D.cand_cap = L.cand_cap; D.node_cap = L.node_cap; D.init_nx = L.init_nx; D.init_ny = L.init_ny;
        D.cand_off = L.cand_off; D.kp_off = L.kp_off;
    }
    // TMA descriptors over the context's own planes (level-0 based ones follow the input: set_level0)
    int mom_w, mom_h, blur_w, blur_h;
    describe_box_dims(&mom_w, &mom_h, &blur_w, &blur_h);
    for (int l = 0; l < p.levels; ++l) {
        Level &L = ctx->lv[l];
        if (int r = encode_plane_map(ctx, &L.map_blur, L.blur, L.w, L.h, L.pitch, L.frame_stride, p.max_frames, blur_w, blur_h, true)) return r;
        if (l > 0)
            if (int r = encode_plane_map(ctx, &L.map_mom, L.pyr, L.w, L.h, L.pitch, L.frame_stride, p.max_frames, mom_w, mom_h)) return r;
    }
    for (int l = 1; l < p.levels; ++l) {
        Level &L = ctx->lv[l];
        if (int r = encode_plane_map(ctx, &L.map_fast, L.pyr, L.w, L.h, L.pitch, L.frame_stride, p.max_frames, 80, 70)) return r;
        if (l + 1 < p.levels && ctx->lv[l + 1].fast_resize)
            if (int r = encode_plane_map(ctx, &ctx->lv[l + 1].map_src, L.pyr, L.w, L.h, L.pitch, L.frame_stride, p.max_frames,
                                         ctx->lv[l + 1].tma_src_w, ctx->lv[l + 1].tma_src_h)) return r;
    }
    if (distribute_smem_bytes(nc_max, false) > 200 * 1024)
        return fail(ctx, SG_ERR_INVALID, "max_keypoints %d needs a quadtree node table larger than shared memory", p.max_keypoints);
    g.cand_per_frame = cand_off;
    g.det_cap = kp_off;
    g.out_cap = kp_off + p.max_tracks;

    {
        std::vector<int4> cells;
        fast_cell_table(g, cells);
        if (int r = dev_alloc(ctx, &ctx->d_cell_table, cells.size())) return r;
        if (!cells.empty()) SG_CUDA(ctx, cudaMemcpy(ctx->d_cell_table, cells.data(), cells.size() * sizeof(int4), cudaMemcpyHostToDevice));
    }
    if (int r = dev_alloc(ctx, &ctx->d_cand, F * cand_off)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_cand_node, F * cand_off)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_cand_count, F * p.levels)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_kp_xy, F * g.det_cap)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_kp_resp, F * g.det_cap)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_kp_count, F * p.levels)) return r;
    // one overflow word for the synchronous calls + one per batch in flight (sg_extract_submit ticket)
    if (int r = dev_alloc(ctx, &ctx->d_err, 1 + sg_ctx::N_TICKETS)) return r;
    SG_CUDA(ctx, cudaMemset(ctx->d_err, 0, sizeof(int) * (1 + sg_ctx::N_TICKETS)));
    SG_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_err, sizeof(int) * (1 + sg_ctx::N_TICKETS), cudaHostAllocDefault));
    for (int i = 0; i <= sg_ctx::N_TICKETS; ++i) ctx->h_err[i] = 0;
    SG_CUDA(ctx, cudaMemset(ctx->d_kp_count, 0, sizeof(int) * F * p.levels));
    const size_t T = std::max(p.max_tracks, 1);
    if (int r = dev_alloc(ctx, &ctx->d_trk_xy, F * T)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_trk_pt, F * T * 2)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_trk_id, F * T)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_trk_count, F)) return r;
    const size_t O = F * g.out_cap;
    if (int r = dev_alloc(ctx, &ctx->d_x, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_y, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_angle, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_octave, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_track_id, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_lvl_x, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_lvl_y, O)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_desc, O * 8 + 8)) return r;   // + one descriptor of slack: a database view may end here
    if (int r = dev_alloc(ctx, &ctx->d_count, F)) return r;
    if (int r = dev_alloc(ctx, &ctx->d_rescans, 1)) return r;
    SG_CUDA(ctx, cudaMemset(ctx->d_rescans, 0, sizeof(unsigned long long)));
    return SG_OK;
}

// A batch still in flight (sg_extract_submit without its sg_extract_wait) owns its frame slots.
static int slots_free(sg_ctx *ctx, int base, int n) {
    for (const auto &o : ctx->ticket)
        if (o.busy && base < o.base + o.n && o.base < base + n)
            return fail(ctx, SG_ERR_INVALID, "frame slots [%d, %d) are still used by a batch in flight ([%d, %d)): wait for it first",
                        base, base + n, o.base, o.base + o.n);
    return SG_OK;
}

static int set_level0(sg_ctx *ctx, const uint8_t *d_imgs, int pitch, size_t stride, int n_frames) {
    const bool changed = ctx->level0 != d_imgs || ctx->level0_pitch != pitch || ctx->level0_stride != stride
                         || ctx->level0_frames < n_frames;
    ctx->level0 = d_imgs; ctx->level0_pitch = pitch; ctx->level0_stride = stride;
    ctx->frames_ready = n_frames;
    ctx->detected = false;
    if (changed) {
        ctx->level0_frames = n_frames;
        return encode_level0_maps(ctx);
    }
    return SG_OK;
}

static int upload_frames(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames) {
    const Level &L0 = ctx->lv[0];
    if (!h_imgs || pitch < L0.w) return fail(ctx, SG_ERR_INVALID, "bad image pointer / pitch");
    if (n_frames < 1 || n_frames > ctx->p.max_frames) return fail(ctx, SG_ERR_INVALID, "n_frames %d outside [1, %d]", n_frames, ctx->p.max_frames);
    for (int f = 0; f < n_frames; ++f)
        SG_CUDA(ctx, cudaMemcpy2DAsync(L0.pyr + (size_t)f * L0.frame_stride, L0.pitch, h_imgs + (size_t)f * frame_stride,
                                       pitch, L0.w, L0.h, cudaMemcpyHostToDevice, ctx->stream));
    return set_level0(ctx, L0.pyr, L0.pitch, L0.frame_stride, n_frames);
}

static int upload_tracks(sg_ctx *ctx, const float *h_xy, const int32_t *h_ids, const int32_t *n_tracks, int n_frames) {
    ctx->have_tracks = false;
    if (!h_xy || !n_tracks || ctx->p.max_tracks <= 0) return SG_OK;
    if (n_frames < 1 || n_frames > ctx->p.max_frames)   // the device arrays hold max_frames * max_tracks entries
        return fail(ctx, SG_ERR_INVALID, "n_frames %d outside [1, %d]", n_frames, ctx->p.max_frames);
    const int T = ctx->p.max_tracks, lvl = ctx->p.track_level;
    if (lvl < 0 || lvl >= ctx->p.levels) return fail(ctx, SG_ERR_INVALID, "track_level outside the pyramid");
    const Level &L = ctx->lv[lvl];
    std::vector<int> xy((size_t)n_frames * T), ids((size_t)n_frames * T), cnt(n_frames);
    std::vector<float> pt((size_t)n_frames * T * 2);
    for (int f = 0; f < n_frames; ++f) {
        if (n_tracks[f] < 0 || n_tracks[f] > T) return fail(ctx, SG_ERR_INVALID, "n_tracks[%d] outside [0, %d]", f, T);
        int n = 0;
        for (int t = 0; t < n_tracks[f]; ++t) {
            // orb_extractor.cpp:89-104: cvRound(pt / scale), 19-px margin (camera validity is the adapter's job)
            const float px = h_xy[2 * ((size_t)f * T + t)], py = h_xy[2 * ((size_t)f * T + t) + 1];
            const int x = (int)lrintf(px / L.scale), y = (int)lrintf(py / L.scale);
            if (x >= PATCH_RADIUS && y >= PATCH_RADIUS && x < L.w - PATCH_RADIUS && y < L.h - PATCH_RADIUS) {
                const size_t o = (size_t)f * T + n;
                xy[o] = x | (y << 16);
                pt[2 * o] = px; pt[2 * o + 1] = py;
                ids[o] = h_ids ? h_ids[(size_t)f * T + t] : t;
                ++n;
            }
        }
        cnt[f] = n;
    }
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->d_trk_xy, xy.data(), xy.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->d_trk_pt, pt.data(), pt.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->d_trk_id, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->d_trk_count, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the staging vectors die here
    ctx->have_tracks = true;
    return SG_OK;
}

static int check_device_images(sg_ctx *ctx, const uint8_t *d_imgs, int pitch, size_t frame_stride, int n_frames) {
    if (!d_imgs || pitch < ctx->lv[0].w || (pitch & 15) || ((uintptr_t)d_imgs & 15) || (frame_stride & 15))
        return fail(ctx, SG_ERR_INVALID, "device images need a 16-byte aligned base, pitch and frame stride");
    if (frame_stride < (size_t)pitch * ctx->lv[0].h) return fail(ctx, SG_ERR_INVALID, "frame_stride smaller than one frame");
    if (n_frames < 1 || n_frames > ctx->p.max_frames) return fail(ctx, SG_ERR_INVALID, "n_frames %d outside [1, %d]", n_frames, ctx->p.max_frames);
    return SG_OK;
}

template <class T>
static int d2h(sg_ctx *ctx, T *h, const T *d, size_t n) {
    if (!h) return SG_OK;
    SG_CUDA(ctx, cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return SG_OK;
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }

int sg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *sg_last_error(const sg_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sg_create(int device, const sg_params *params, sg_ctx **out) {
    if (!params || !out) return fail(nullptr, SG_ERR_INVALID, "null argument");
    *out = nullptr;
    const sg_params &p = *params;
    if (p.width < 64 || p.height < 64 || p.levels < 1 || p.levels > SG_MAX_LEVELS || !(p.scale_factor > 1.0f)
        || p.max_keypoints < 1 || p.max_frames < 1 || p.max_tracks < 0 || p.min_fast_thr < 1
        || p.ini_fast_thr < p.min_fast_thr || p.ini_fast_thr > 254
        || (p.max_tracks > 0 && (p.track_level < 0 || p.track_level >= p.levels)))
        return fail(nullptr, SG_ERR_INVALID, "invalid sg_params");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, SG_ERR_CUDA, "no CUDA device: %s (libslamgpu has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, SG_ERR_INVALID, "device %d outside [0, %d)", device, n);
    sg_ctx *ctx = new sg_ctx();
    ctx->device = device;
    ctx->p = p;
    auto bail = [&](int code) { g_create_error = ctx->err; sg_destroy(ctx); return code; };
    if (cudaSetDevice(device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(SG_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return bail(SG_ERR_CUDA); }
    if (prop.major < 10) { ctx->err = "libslamgpu is built for sm_100a (B200) only"; return bail(SG_ERR_CUDA); }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        ctx->err = "stream / event creation failed";
        return bail(SG_ERR_CUDA);
    }
    ctx->main_stream = ctx->stream;
    {   // the context's own stream-ordered memory pool (scratch buffers, descriptor databases): it keeps what is freed
        // -- no trimming at synchronisation points -- and leaves the process-wide default pool alone
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess) { ctx->err = "cudaMemPoolCreate failed"; return bail(SG_ERR_CUDA); }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (const char *e = getenv("SG_PIPE_STREAMS")) ctx->pipe_streams = std::min((int)sg_ctx::N_CMP, std::max(1, atoi(e)));   // tuning knob
    if (cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        ctx->err = "stream / event creation failed";
        return bail(SG_ERR_CUDA);
    }
    for (auto &q : ctx->s_cmp)
        if (cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "stream creation failed"; return bail(SG_ERR_CUDA); }
    for (auto &slot : ctx->ev_stage)
        for (auto &e : slot)
            if (cudaEventCreate(&e) != cudaSuccess) { ctx->err = "event creation failed"; return bail(SG_ERR_CUDA); }
    if (int r = build_context(ctx)) return bail(r);
    *out = ctx;
    return SG_OK;
}

void sg_destroy(sg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    // batches may still be in flight on the pipeline streams (sg_extract_submit without its wait)
    for (cudaStream_t q : {ctx->main_stream, ctx->s_in, ctx->s_out}) if (q) cudaStreamSynchronize(q);
    for (cudaStream_t q : ctx->s_cmp) if (q) cudaStreamSynchronize(q);
    // databases still alive: their device memory goes with the context (pool); the handles stay valid for sg_db_destroy
    for (sg_db *db : ctx->dbs) {
        if (db->owns_data) { if (db->d_desc) cudaFree(db->d_desc); if (db->d_angle) cudaFree(db->d_angle); }
        if (db->d_offsets) cudaFree(db->d_offsets);
        db->d_desc = nullptr; db->d_angle = nullptr; db->d_offsets = nullptr; db->ctx = nullptr;
    }
    ctx->dbs.clear();
    for (auto &L : ctx->lv) { cudaFree(L.pyr); cudaFree(L.blur); cudaFree(L.xtab); cudaFree(L.ytab); cudaFree(L.xtile); cudaFree(L.ytile); }
    void *ptrs[] = {ctx->d_cand, ctx->d_cand_node, ctx->d_cand_count, ctx->d_kp_xy, ctx->d_kp_resp, ctx->d_kp_count,
                    ctx->d_err, ctx->d_trk_xy, ctx->d_trk_pt, ctx->d_trk_id, ctx->d_trk_count, ctx->d_x, ctx->d_y,
                    ctx->d_angle, ctx->d_octave, ctx->d_track_id, ctx->d_lvl_x, ctx->d_lvl_y, ctx->d_desc, ctx->d_count,
                    ctx->d_flush, ctx->d_topk, ctx->d_nseen, ctx->d_pairs, ctx->d_matches, ctx->d_nmatch,
                    ctx->d_rescans, ctx->d_tmp, ctx->d_dbtmp, ctx->d_cell_table};
    for (void *q : ptrs) if (q) cudaFree(q);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->h_pack) cudaFreeHost(ctx->h_pack);
    if (ctx->d_pack) cudaFree(ctx->d_pack);
    for (auto &slot : ctx->ev_stage)
        for (auto &e : slot) if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (auto &e : ctx->pipe_ev) cudaEventDestroy(e);
    for (auto &e : ctx->ticket_ev) if (e) cudaEventDestroy(e);
    for (cudaStream_t q : {ctx->s_in, ctx->s_out}) if (q) cudaStreamDestroy(q);
    for (cudaStream_t q : ctx->s_cmp) if (q) cudaStreamDestroy(q);
    if (cudaStream_t m = ctx->main_stream ? ctx->main_stream : ctx->stream) cudaStreamDestroy(m);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

int sg_synchronize(sg_ctx *ctx) {
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}
void *sg_stream(sg_ctx *ctx) { return (void *)ctx->stream; }
unsigned long long sg_launch_count(const sg_ctx *ctx) { return ctx->launches; }

int sg_get_geometry(const sg_ctx *ctx, float *scales, int *widths, int *heights, int *pitches, int *budgets) {
    for (int l = 0; l < ctx->p.levels; ++l) {
        const Level &L = ctx->lv[l];
        if (scales) scales[l] = L.scale;
        if (widths) widths[l] = L.w;
        if (heights) heights[l] = L.h;
        if (pitches) pitches[l] = L.pitch;
        if (budgets) budgets[l] = L.budget;
    }
    return SG_OK;
}
int sg_keypoint_capacity(const sg_ctx *ctx) { return ctx->geom.out_cap; }

// ---- pyramid ------------------------------------------------------------------------------------------
int sg_pyramid_update(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames) {
    cudaSetDevice(ctx->device);
    if (int r = slots_free(ctx, 0, n_frames)) return r;
    if (int r = upload_frames(ctx, h_imgs, pitch, frame_stride, n_frames)) return r;
    if (int r = launch_pyramid(ctx, n_frames)) return r;
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

int sg_pyramid_update_device(sg_ctx *ctx, const uint8_t *d_imgs, int pitch, size_t frame_stride, int n_frames) {
    cudaSetDevice(ctx->device);
    if (int r = slots_free(ctx, 0, n_frames)) return r;
    if (int r = check_device_images(ctx, d_imgs, pitch, frame_stride, n_frames)) return r;
    if (int r = set_level0(ctx, d_imgs, pitch, frame_stride, n_frames)) return r;
    return launch_pyramid(ctx, n_frames);
}

int sg_pyramid_download(sg_ctx *ctx, int frame, int level, int blurred, uint8_t *h_dst, int dst_pitch) {
    cudaSetDevice(ctx->device);
    if (frame < 0 || frame >= ctx->frames_ready || level < 0 || level >= ctx->p.levels || !h_dst)
        return fail(ctx, SG_ERR_INVALID, "bad frame / level");
    const Level &L = ctx->lv[level];
    const uint8_t *src;
    int pitch;
    if (blurred) { src = L.blur + (size_t)frame * L.frame_stride; pitch = L.pitch; }
    else if (level == 0) { src = ctx->level0 + (size_t)frame * ctx->level0_stride; pitch = ctx->level0_pitch; }
    else { src = L.pyr + (size_t)frame * L.frame_stride; pitch = L.pitch; }
    SG_CUDA(ctx, cudaMemcpy2DAsync(h_dst, dst_pitch, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

int sg_pyramid_device_plane(sg_ctx *ctx, int level, int blurred, const uint8_t **d_plane, int *pitch, size_t *frame_stride) {
    if (level < 0 || level >= ctx->p.levels) return fail(ctx, SG_ERR_INVALID, "bad level");
    const Level &L = ctx->lv[level];
    if (!blurred && level == 0) {
        if (d_plane) *d_plane = ctx->level0;
        if (pitch) *pitch = ctx->level0_pitch;
        if (frame_stride) *frame_stride = ctx->level0_stride;
    } else {
        if (d_plane) *d_plane = blurred ? L.blur : L.pyr;
        if (pitch) *pitch = L.pitch;
        if (frame_stride) *frame_stride = L.frame_stride;
    }
    return SG_OK;
}

// ---- detection ----------------------------------------------------------------------------------------
int sg_detect(sg_ctx *ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->frames_ready < 1) return fail(ctx, SG_ERR_STATE, "sg_detect before sg_pyramid_update");
    if (int r = launch_detect(ctx, ctx->frames_ready)) return r;
    return check_device_error(ctx);
}

int sg_detect_download(sg_ctx *ctx, int frame, int level, int *h_x, int *h_y, int *h_resp, int cap, int *n) {
    cudaSetDevice(ctx->device);
    if (!ctx->detected) return fail(ctx, SG_ERR_STATE, "sg_detect_download before sg_detect");
    if (frame < 0 || frame >= ctx->frames_ready || level < 0 || level >= ctx->p.levels) return fail(ctx, SG_ERR_INVALID, "bad frame / level");
    const GeomDev &g = ctx->geom;
    int cnt = 0;
    SG_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->d_kp_count + frame * g.levels + level, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n) *n = cnt;
    const int m = std::min(cnt, cap);
    if (m <= 0) return SG_OK;
    std::vector<int> xy(m), rs(m);
    const size_t off = (size_t)frame * g.det_cap + g.lv[level].kp_off;
    SG_CUDA(ctx, cudaMemcpyAsync(xy.data(), ctx->d_kp_xy + off, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(rs.data(), ctx->d_kp_resp + off, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < m; ++i) {
        if (h_x) h_x[i] = xy[i] & 0xffff;
        if (h_y) h_y[i] = xy[i] >> 16;
        if (h_resp) h_resp[i] = rs[i];
    }
    return SG_OK;
}

int sg_detect_download_candidates(sg_ctx *ctx, int frame, int level, int *h_x, int *h_y, int *h_resp, int cap, int *n) {
    cudaSetDevice(ctx->device);
    if (!ctx->detected) return fail(ctx, SG_ERR_STATE, "no detection yet");
    if (frame < 0 || frame >= ctx->frames_ready || level < 0 || level >= ctx->p.levels) return fail(ctx, SG_ERR_INVALID, "bad frame / level");
    const GeomDev &g = ctx->geom;
    int cnt = 0;
    SG_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->d_cand_count + frame * g.levels + level, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n) *n = cnt;
    const int m = std::min(cnt, cap);
    if (m <= 0) return SG_OK;
    std::vector<unsigned long long> c(m);
    SG_CUDA(ctx, cudaMemcpyAsync(c.data(), ctx->d_cand + (size_t)frame * g.cand_per_frame + g.lv[level].cand_off,
                                 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < m; ++i) {
        const unsigned key = (unsigned)c[i];
        if (h_x) h_x[i] = FAST_BORDER + CELL * ((key >> 12) & 1023) + (key & 63);
        if (h_y) h_y[i] = FAST_BORDER + CELL * (key >> 22) + ((key >> 6) & 63);
        if (h_resp) h_resp[i] = (int)(c[i] >> 32);
    }
    return SG_OK;
}

// ---- extraction ---------------------------------------------------------------------------------------
static int extract_launches(sg_ctx *ctx, int n_frames) {
    if (int r = launch_pyramid(ctx, n_frames)) return r;
    if (int r = launch_detect(ctx, n_frames)) return r;
    return launch_describe(ctx, n_frames);
}

int sg_extract_download(sg_ctx *ctx, int n_frames, sg_keypoints *o) {
    cudaSetDevice(ctx->device);
    if (!o || n_frames < 1 || n_frames > ctx->frames_ready) return fail(ctx, SG_ERR_INVALID, "bad n_frames / output");
    if (int r = slots_free(ctx, 0, n_frames)) return r;     // a batch in flight is still writing these slots
    const size_t n = (size_t)n_frames * ctx->geom.out_cap;
    if (int r = d2h(ctx, o->x, ctx->d_x, n)) return r;
    if (int r = d2h(ctx, o->y, ctx->d_y, n)) return r;
    if (int r = d2h(ctx, o->angle, ctx->d_angle, n)) return r;
    if (int r = d2h(ctx, o->octave, ctx->d_octave, n)) return r;
    if (int r = d2h(ctx, o->desc, ctx->d_desc, n * 8)) return r;
    if (int r = d2h(ctx, o->track_id, ctx->d_track_id, n)) return r;
    if (int r = d2h(ctx, o->lvl_x, ctx->d_lvl_x, n)) return r;
    if (int r = d2h(ctx, o->lvl_y, ctx->d_lvl_y, n)) return r;
    if (int r = d2h(ctx, o->count, ctx->d_count, (size_t)n_frames)) return r;
    if (int r = d2h(ctx, o->level_count, ctx->d_kp_count, (size_t)n_frames * ctx->p.levels)) return r;
    return check_device_error(ctx);
}

// Host-buffer extraction as a three-stage pipeline over chunks of the batch: H2D of chunk c+1 (copy engine,
// stream s_in) runs under the kernels of chunk c (alternating compute streams, so the tail of one chunk's
// small-level kernels overlaps the head of the next) and the D2H of chunk c-1 (second copy engine, s_out).
// Frames are independent, every per-frame buffer is indexed by the absolute frame number, so the chunks
// never touch the same memory.
// Queues the whole batch (H2D, kernels, D2H) on the pipeline streams; frames occupy the context's frame slots
// [base, base + n_frames), host arrays are indexed from 0.  Nothing is synchronised here.
static int pipeline_submit(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames, int base,
                           const sg_keypoints *o, bool streaming = false) {
    const Level &L0 = ctx->lv[0];
    if (!h_imgs || pitch < L0.w) return fail(ctx, SG_ERR_INVALID, "bad image pointer / pitch");
    if (n_frames < 1 || base < 0 || base + n_frames > ctx->p.max_frames)
        return fail(ctx, SG_ERR_INVALID, "frames [%d, %d) outside the context's [0, %d)", base, base + n_frames, ctx->p.max_frames);
    if (!o) return fail(ctx, SG_ERR_INVALID, "null output");
    // chunk schedule: small chunks at both ends (short pipeline fill and drain: the first kernels start after a
    // short copy, the last copy-out is short), full chunks in between (efficient grids)
    std::vector<int> sched;
    {
        // (a stream of batches keeps the pipeline full by itself: uniform, larger chunks)
        const int C = std::max(1, streaming ? ctx->stream_chunk : ctx->pipe_chunk), q = std::max(1, C / 4), h = std::max(1, C / 2);
        int rest = n_frames;
        if (!streaming && n_frames >= 2 * (q + h) + C) {
            sched.push_back(q); sched.push_back(h);
            rest -= 2 * (q + h);
            while (rest > 0) { const int n = std::min(C, rest); sched.push_back(n); rest -= n; }
            sched.push_back(h); sched.push_back(q);
        } else {
            while (rest > 0) { const int n = std::min(C, rest); sched.push_back(n); rest -= n; }
        }
    }
    const int chunks = (int)sched.size();
    while ((int)ctx->pipe_ev.size() < 2 * chunks) {
        cudaEvent_t e;
        SG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    if (int r = set_level0(ctx, L0.pyr, L0.pitch, L0.frame_stride, base + n_frames)) return r;
    // work queued earlier on the main stream (an un-synchronised sg_extract_device, ...) comes first
    SG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->main_stream));
    for (cudaStream_t q : {ctx->s_in, ctx->s_out}) SG_CUDA(ctx, cudaStreamWaitEvent(q, ctx->ev_fork, 0));
    for (cudaStream_t q : ctx->s_cmp) SG_CUDA(ctx, cudaStreamWaitEvent(q, ctx->ev_fork, 0));
    const size_t cap = ctx->geom.out_cap;
    const int levels = ctx->p.levels;
    int rc = SG_OK;
    int f_next = 0;
    static const bool dbg = getenv("SG_DEBUG_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    for (int c = 0; c < chunks && rc == SG_OK; ++c) {
        const int fl = f_next, f0 = base + f_next, n = sched[c];     // fl: index in the host arrays, f0: frame slot
        f_next += n;
        if (pitch == L0.pitch && (n == 1 || frame_stride == L0.frame_stride)) {
            SG_CUDA(ctx, cudaMemcpyAsync(L0.pyr + (size_t)f0 * L0.frame_stride, h_imgs + (size_t)fl * frame_stride,
                                         (size_t)n * L0.frame_stride, cudaMemcpyHostToDevice, ctx->s_in));
        } else {
            for (int f = 0; f < n; ++f)
                SG_CUDA(ctx, cudaMemcpy2DAsync(L0.pyr + (size_t)(f0 + f) * L0.frame_stride, L0.pitch, h_imgs + (size_t)(fl + f) * frame_stride,
                                               pitch, L0.w, L0.h, cudaMemcpyHostToDevice, ctx->s_in));
        }
        SG_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[2 * c], ctx->s_in));
        cudaStream_t cmp = ctx->s_cmp[ctx->pipe_rr++ % (unsigned)ctx->pipe_streams];   // rotates across calls too (batches in flight)
        SG_CUDA(ctx, cudaStreamWaitEvent(cmp, ctx->pipe_ev[2 * c], 0));
        ctx->stream = cmp; ctx->frame0 = f0; ctx->in_pipeline = true;
        rc = extract_launches(ctx, n);
        ctx->stream = ctx->main_stream; ctx->frame0 = 0; ctx->in_pipeline = false;
        if (rc) break;
        SG_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[2 * c + 1], cmp));
        SG_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->pipe_ev[2 * c + 1], 0));
        auto out = [&](auto *h, const auto *d, size_t per_frame) -> int {
            if (!h) return SG_OK;
            SG_CUDA(ctx, cudaMemcpyAsync(h + (size_t)fl * per_frame, d + (size_t)f0 * per_frame,
                                         (size_t)n * per_frame * sizeof(*h), cudaMemcpyDeviceToHost, ctx->s_out));
            return SG_OK;
        };
        if ((rc = out(o->x, ctx->d_x, cap)) || (rc = out(o->y, ctx->d_y, cap)) || (rc = out(o->angle, ctx->d_angle, cap))
            || (rc = out(o->octave, ctx->d_octave, cap)) || (rc = out(o->desc, ctx->d_desc, 8 * cap))
            || (rc = out(o->track_id, ctx->d_track_id, cap)) || (rc = out(o->lvl_x, ctx->d_lvl_x, cap))
            || (rc = out(o->lvl_y, ctx->d_lvl_y, cap)) || (rc = out(o->count, ctx->d_count, 1))
            || (rc = out(o->level_count, ctx->d_kp_count, (size_t)levels)))
            break;
    }
    if (dbg)
        fprintf(stderr, "sg_extract: %d chunks, submit %.3f ms\n", chunks,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    return rc;
}

static int pipeline_wait_all(sg_ctx *ctx, int rc) {
    for (int i = 0; i <= sg_ctx::N_CMP; ++i) {
        const cudaError_t e = cudaStreamSynchronize(i < sg_ctx::N_CMP ? ctx->s_cmp[i] : ctx->s_out);
        if (e != cudaSuccess && rc == SG_OK) rc = fail(ctx, SG_ERR_CUDA, "pipeline synchronise failed: %s", cudaGetErrorString(e));
    }
    if (rc) return rc;
    return check_device_error(ctx);
}

// ---- small batches (the reference's own call pattern: one keyframe per call, keyframe.cpp:95-116) ---------------------
// Latency path: one stream, and ONE device->host copy.  A pack kernel gathers what the call returns for a frame --
// keypoint count, per-level counts, the overflow word and the `count` live entries of every output array -- into one
// block of a staging buffer; the host scatters the block into the caller's arrays.  (The pipelined path issues up to
// ten copies per chunk plus a separate read of the overflow word: ~35 us of a ~200 us call.)
constexpr int LEAN_MAX_FRAMES = 4;
constexpr int PACK_HEADER_WORDS = 32;    // count, err, level counts (<= SG_MAX_LEVELS), padding: 128 bytes

__global__ void __launch_bounds__(256)
pack_outputs_kernel(const float *x, const float *y, const float *angle, const int *octave, const int *track_id, const int *lvl_x,
                    const int *lvl_y, const uint32_t *desc, const int *count, const int *kp_count, const int *err, int levels,
                    int out_cap, int cap, uint32_t *pack) {
    // out_cap: slots per frame of the source arrays; cap: out_cap rounded up to 4 (16-byte aligned fields in the block)
    const int f = blockIdx.x;
    const size_t block_words = PACK_HEADER_WORDS + (size_t)cap * 15;
    uint32_t *out = pack + (size_t)f * block_words;
    const int n = min(count[f], out_cap);
    if (threadIdx.x == 0) { out[0] = (uint32_t)count[f]; out[1] = (uint32_t)*err; }
    if ((int)threadIdx.x < levels) out[2 + threadIdx.x] = (uint32_t)kp_count[f * levels + threadIdx.x];
    uint32_t *o = out + PACK_HEADER_WORDS;
    const size_t base = (size_t)f * out_cap;
    const uint32_t *src[7] = {reinterpret_cast<const uint32_t *>(x), reinterpret_cast<const uint32_t *>(y),
                              reinterpret_cast<const uint32_t *>(angle), reinterpret_cast<const uint32_t *>(octave),
                              reinterpret_cast<const uint32_t *>(track_id), reinterpret_cast<const uint32_t *>(lvl_x),
                              reinterpret_cast<const uint32_t *>(lvl_y)};
#pragma unroll
    for (int a = 0; a < 7; ++a)
        for (int i = threadIdx.x; i < n; i += blockDim.x) o[(size_t)a * cap + i] = src[a][base + i];
    const uint4 *d4 = reinterpret_cast<const uint4 *>(desc + 8 * base);
    uint4 *o4 = reinterpret_cast<uint4 *>(o + (size_t)7 * cap);
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) o4[i] = d4[i];
}

static int extract_lean(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames, sg_keypoints *o) {
    const Level &L0 = ctx->lv[0];
    if (!h_imgs || pitch < L0.w) return fail(ctx, SG_ERR_INVALID, "bad image pointer / pitch");
    if (!o) return fail(ctx, SG_ERR_INVALID, "null output");
    const int out_cap = ctx->geom.out_cap, cap = (out_cap + 3) & ~3, levels = ctx->p.levels;
    const size_t block_words = PACK_HEADER_WORDS + (size_t)cap * 15, bytes = (size_t)n_frames * block_words * 4;
    if (!ctx->d_pack) {
        SG_CUDA(ctx, cudaMalloc((void **)&ctx->d_pack, (size_t)LEAN_MAX_FRAMES * block_words * 4));
        SG_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_pack, (size_t)LEAN_MAX_FRAMES * block_words * 4, cudaHostAllocDefault));
    }
    cudaStream_t st = ctx->main_stream;
    if (pitch == L0.pitch && (n_frames == 1 || frame_stride == L0.frame_stride)) {
        SG_CUDA(ctx, cudaMemcpyAsync(L0.pyr, h_imgs, (size_t)(n_frames - 1) * L0.frame_stride + (size_t)L0.pitch * L0.h,
                                     cudaMemcpyHostToDevice, st));
    } else {
        for (int f = 0; f < n_frames; ++f)
            SG_CUDA(ctx, cudaMemcpy2DAsync(L0.pyr + (size_t)f * L0.frame_stride, L0.pitch, h_imgs + (size_t)f * frame_stride, pitch,
                                           L0.w, L0.h, cudaMemcpyHostToDevice, st));
    }
    if (int r = set_level0(ctx, L0.pyr, L0.pitch, L0.frame_stride, n_frames)) return r;
    ctx->stream = st; ctx->frame0 = 0;
    if (int r = extract_launches(ctx, n_frames)) return r;
    pack_outputs_kernel<<<n_frames, 256, 0, st>>>(ctx->d_x, ctx->d_y, ctx->d_angle, ctx->d_octave, ctx->d_track_id, ctx->d_lvl_x,
                                                  ctx->d_lvl_y, ctx->d_desc, ctx->d_count, ctx->d_kp_count, ctx->d_err, levels, out_cap,
                                                  cap, ctx->d_pack);
    SG_LAUNCH_CHECK(ctx);
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->h_pack, ctx->d_pack, bytes, cudaMemcpyDeviceToHost, st));
    SG_CUDA(ctx, cudaStreamSynchronize(st));
    int err = 0;
    for (int f = 0; f < n_frames; ++f) {
        const uint32_t *b = ctx->h_pack + (size_t)f * block_words, *a = b + PACK_HEADER_WORDS;
        const int n = std::min((int)b[0], out_cap);
        err |= (int)b[1];
        if (o->count) o->count[f] = (int)b[0];
        if (o->level_count) std::memcpy(o->level_count + (size_t)f * levels, b + 2, sizeof(int) * (size_t)levels);
        const size_t at = (size_t)f * out_cap;
        auto put = [&](auto *dst, int field, size_t words_per) {
            if (dst) std::memcpy(dst + at * words_per, a + (size_t)field * cap, (size_t)n * words_per * 4);
        };
        put(o->x, 0, 1); put(o->y, 1, 1); put(o->angle, 2, 1); put(o->octave, 3, 1); put(o->track_id, 4, 1);
        put(o->lvl_x, 5, 1); put(o->lvl_y, 6, 1); put(o->desc, 7, 8);
    }
    if (err) {
        cudaMemsetAsync(ctx->d_err, 0, sizeof(int), st);
        return fail(ctx, err, "device-side capacity exceeded (candidate list or quadtree node table)");
    }
    return SG_OK;
}

int sg_extract(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames,
               const float *h_track_xy, const int32_t *h_track_ids, const int32_t *n_tracks, sg_keypoints *h_out) {
    cudaSetDevice(ctx->device);
    if (n_frames < 1 || n_frames > ctx->p.max_frames)   // before anything is sized by it (track staging, device copies)
        return fail(ctx, SG_ERR_INVALID, "n_frames %d outside [1, %d]", n_frames, ctx->p.max_frames);
    if (int r = slots_free(ctx, 0, n_frames)) return r;
    if (int r = upload_tracks(ctx, h_track_xy, h_track_ids, n_tracks, n_frames)) return r;
    if (n_frames <= LEAN_MAX_FRAMES && !ctx->profiling) return extract_lean(ctx, h_imgs, pitch, frame_stride, n_frames, h_out);
    return pipeline_wait_all(ctx, pipeline_submit(ctx, h_imgs, pitch, frame_stride, n_frames, 0, h_out));
}

// Streaming form: several batches in flight on disjoint frame slots of the context (see slamgpu.h).
int sg_extract_submit(sg_ctx *ctx, const uint8_t *h_imgs, int pitch, size_t frame_stride, int n_frames, int base_frame,
                      const sg_keypoints *h_out, int *ticket) {
    cudaSetDevice(ctx->device);
    if (!ticket) return fail(ctx, SG_ERR_INVALID, "null ticket");
    if (n_frames < 1 || n_frames > ctx->p.max_frames)
        return fail(ctx, SG_ERR_INVALID, "n_frames %d outside [1, %d]", n_frames, ctx->p.max_frames);
    int t = -1;                                   // any free ticket: waits may come back out of order
    for (int i = 0; i < sg_ctx::N_TICKETS && t < 0; ++i) {
        const int c = (ctx->next_ticket + i) % sg_ctx::N_TICKETS;
        if (!ctx->ticket[c].busy) t = c;
    }
    if (t < 0) return fail(ctx, SG_ERR_INVALID, "%d batches are in flight: wait for one first", (int)sg_ctx::N_TICKETS);
    if (int r = slots_free(ctx, base_frame, n_frames)) return r;
    ctx->have_tracks = false;
    ctx->err_slot = 1 + t;                        // this batch's own overflow word
    const int rs = pipeline_submit(ctx, h_imgs, pitch, frame_stride, n_frames, base_frame, h_out, true);
    ctx->err_slot = 0;
    if (rs) {
        pipeline_wait_all(ctx, rs);
        return rs;
    }
    ctx->next_ticket = t + 1;
    ctx->ticket[t].busy = true; ctx->ticket[t].base = base_frame; ctx->ticket[t].n = n_frames;
    if (!ctx->ticket_ev[t]) SG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ticket_ev[t], cudaEventDisableTiming));
    // s_out runs behind every kernel of the batch: read the batch's overflow word there and clear it for the next user
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->h_err + 1 + t, ctx->d_err + 1 + t, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_out));
    SG_CUDA(ctx, cudaMemsetAsync(ctx->d_err + 1 + t, 0, sizeof(int), ctx->s_out));
    SG_CUDA(ctx, cudaEventRecord(ctx->ticket_ev[t], ctx->s_out));      // every D2H of the batch is on s_out, after its kernels
    *ticket = t;
    return SG_OK;
}

int sg_extract_wait(sg_ctx *ctx, int ticket) {
    cudaSetDevice(ctx->device);
    if (ticket < 0 || ticket >= sg_ctx::N_TICKETS || !ctx->ticket[ticket].busy) return fail(ctx, SG_ERR_INVALID, "ticket %d is not in flight", ticket);
    SG_CUDA(ctx, cudaEventSynchronize(ctx->ticket_ev[ticket]));
    ctx->ticket[ticket].busy = false;
    const int e = ctx->h_err[1 + ticket];
    ctx->h_err[1 + ticket] = 0;
    if (e) return fail(ctx, e, "device-side capacity exceeded (candidate list or quadtree node table) in the batch of ticket %d", ticket);
    return SG_OK;
}

int sg_set_pipeline_chunk(sg_ctx *ctx, int frames) {
    if (frames < 1) return fail(ctx, SG_ERR_INVALID, "chunk must be >= 1 frame");
    ctx->pipe_chunk = frames;
    ctx->stream_chunk = frames;
    return SG_OK;
}

int sg_extract_device(sg_ctx *ctx, const uint8_t *d_imgs, int pitch, size_t frame_stride, int n_frames) {
    cudaSetDevice(ctx->device);
    if (int r = slots_free(ctx, 0, n_frames)) return r;
    if (int r = check_device_images(ctx, d_imgs, pitch, frame_stride, n_frames)) return r;
    if (int r = set_level0(ctx, d_imgs, pitch, frame_stride, n_frames)) return r;
    ctx->have_tracks = false;
    // (more slices than streams: the slices rotate over the streams; small slices keep a slice's planes inside the L2)
    const int parts = std::max(1, std::min(ctx->overlap_parts, n_frames / 8));
    if (parts <= 1 || ctx->profiling) return extract_launches(ctx, n_frames);
    // Independent slices of the batch on separate streams: the latency-bound quadtree kernel and the kernel tails
    // of one slice run under the issue-bound pyramid / FAST kernels of another.  Joined back on the main stream.
    const int streams = std::min(parts, ctx->overlap_streams);
    while ((int)ctx->pipe_ev.size() < streams) {
        cudaEvent_t e;
        SG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    SG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->main_stream));
    int rc = SG_OK, f0 = 0;
    for (int c = 0; c < parts && rc == SG_OK; ++c) {
        const int n = (n_frames - f0) / (parts - c);
        cudaStream_t cmp = ctx->s_cmp[c % streams];
        if (c < streams) SG_CUDA(ctx, cudaStreamWaitEvent(cmp, ctx->ev_fork, 0));
        ctx->stream = cmp; ctx->frame0 = f0; ctx->in_pipeline = true;
        rc = extract_launches(ctx, n);
        ctx->stream = ctx->main_stream; ctx->frame0 = 0; ctx->in_pipeline = false;
        if (rc) break;
        f0 += n;
    }
    for (int c = 0; c < streams && rc == SG_OK; ++c) {     // join every stream back on the main stream
        SG_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[c], ctx->s_cmp[c]));
        SG_CUDA(ctx, cudaStreamWaitEvent(ctx->main_stream, ctx->pipe_ev[c], 0));
    }
    return rc;
}

int sg_set_overlap(sg_ctx *ctx, int parts) {
    // parts: slices of the batch (1..64); values above 1000 encode "streams * 1000 + parts" (tuning aid)
    int streams = ctx->overlap_streams;
    if (parts >= 1000) { streams = parts / 1000; parts %= 1000; }
    if (parts < 1 || parts > 64 || streams < 1 || streams > sg_ctx::N_CMP)
        return fail(ctx, SG_ERR_INVALID, "parts must be 1..64 and streams 1..%d", (int)sg_ctx::N_CMP);
    ctx->overlap_parts = parts;
    ctx->overlap_streams = streams;
    return SG_OK;
}

int sg_extract_device_views(sg_ctx *ctx, sg_keypoints_dev *out) {
    if (!out) return fail(ctx, SG_ERR_INVALID, "null argument");
    out->x = ctx->d_x; out->y = ctx->d_y; out->angle = ctx->d_angle; out->octave = ctx->d_octave;
    out->desc = ctx->d_desc; out->count = ctx->d_count; out->cap = ctx->geom.out_cap;
    return SG_OK;
}

// ---- matching -----------------------------------------------------------------------------------------
int sg_hamming(sg_ctx *ctx, const uint32_t *h_a, const uint32_t *h_b, int n, uint32_t *h_out) {
    cudaSetDevice(ctx->device);
    if (n <= 0) return SG_OK;
    if (!h_a || !h_b || !h_out) return fail(ctx, SG_ERR_INVALID, "null argument");
    const size_t bytes = (size_t)n * 32;
    if (int r = grow(ctx, &ctx->d_tmp, &ctx->tmp_bytes, 2 * bytes + (size_t)n * 4, 1)) return r;
    uint8_t *base = (uint8_t *)ctx->d_tmp;
    SG_CUDA(ctx, cudaMemcpyAsync(base, h_a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(base + bytes, h_b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (int r = run_hamming(ctx, (uint32_t *)base, (uint32_t *)(base + bytes), n, (uint32_t *)(base + 2 * bytes))) return r;
    SG_CUDA(ctx, cudaMemcpyAsync(h_out, base + 2 * bytes, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

static int db_finish(sg_ctx *ctx, sg_db *db, const int64_t *h_offsets, int n_sets) {
    db->ctx = ctx;
    ctx->dbs.push_back(db);      // sg_destroy releases the device memory of databases that outlive their context
    db->n_sets = n_sets;
    db->offsets.assign(h_offsets, h_offsets + n_sets + 1);
    db->max_set = 0;
    for (int s = 0; s < n_sets; ++s) {
        if (h_offsets[s + 1] < h_offsets[s]) return fail(ctx, SG_ERR_INVALID, "offsets must be non-decreasing");
        db->max_set = std::max<long long>(db->max_set, h_offsets[s + 1] - h_offsets[s]);
    }
    // stream-ordered allocation from the device's memory pool: creating / destroying a database (per step in the
    // extract -> match flow) must not pay a device-wide cudaMalloc / cudaFree
    SG_CUDA(ctx, cudaMallocFromPoolAsync((void **)&db->d_offsets, sizeof(long long) * (n_sets + 1), ctx->pool, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(db->d_offsets, db->offsets.data(), sizeof(long long) * (n_sets + 1), cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // db->offsets is pageable host memory that outlives the call, but keep the create synchronous
    return SG_OK;
}

static int db_create(sg_ctx *ctx, const uint32_t *desc, const float *angle, const int64_t *h_offsets, int n_sets,
                     sg_db **out, cudaMemcpyKind kind) {
    cudaSetDevice(ctx->device);
    if (!desc || !angle || !h_offsets || n_sets < 1 || !out) return fail(ctx, SG_ERR_INVALID, "null / empty database");
    if (h_offsets[0] != 0) return fail(ctx, SG_ERR_INVALID, "offsets[0] must be 0");
    sg_db *db = new sg_db();
    const size_t total = (size_t)h_offsets[n_sets];
    int r = db_finish(ctx, db, h_offsets, n_sets);
    if (!r && cudaMallocFromPoolAsync((void **)&db->d_desc, std::max<size_t>(total, 1) * 32 + 32, ctx->pool, ctx->stream) != cudaSuccess) r = fail(ctx, SG_ERR_CUDA, "pool allocation failed");
    if (!r && cudaMallocFromPoolAsync((void **)&db->d_angle, std::max<size_t>(total, 1) * 4, ctx->pool, ctx->stream) != cudaSuccess) r = fail(ctx, SG_ERR_CUDA, "pool allocation failed");
    if (!r && total) {
        if (cudaMemcpyAsync(db->d_desc, desc, total * 32, kind, ctx->stream) != cudaSuccess
            || cudaMemcpyAsync(db->d_angle, angle, total * 4, kind, ctx->stream) != cudaSuccess
            || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            r = fail(ctx, SG_ERR_CUDA, "database upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (r) { sg_db_destroy(db); return r; }
    *out = db;
    return SG_OK;
}

// A transient database for the single-call matchers (sg_match_bruteforce / _bow / _triangulation): descriptors, angles
// and offsets live in a grow-only buffer of the context, so a call costs no cudaMalloc / cudaFree.
static int scratch_db(sg_ctx *ctx, sg_db *db, const uint32_t *h_desc, const float *h_angle, const int64_t *h_offsets, int n_sets) {
    db->ctx = ctx;
    db->n_sets = n_sets;
    db->offsets.assign(h_offsets, h_offsets + n_sets + 1);
    db->max_set = 0;
    for (int s = 0; s < n_sets; ++s) db->max_set = std::max<long long>(db->max_set, h_offsets[s + 1] - h_offsets[s]);
    const size_t total = (size_t)h_offsets[n_sets];
    const size_t b_desc = (total * 32 + 32 + 255) & ~(size_t)255, b_ang = (total * 4 + 4 + 255) & ~(size_t)255;
    const size_t b_off = (sizeof(long long) * ((size_t)n_sets + 1) + 255) & ~(size_t)255;
    if (int r = grow(ctx, &ctx->d_dbtmp, &ctx->dbtmp_bytes, b_desc + b_ang + b_off, 1)) return r;
    uint8_t *base = (uint8_t *)ctx->d_dbtmp;
    db->d_desc = (uint32_t *)base;
    db->d_angle = (float *)(base + b_desc);
    db->d_offsets = (long long *)(base + b_desc + b_ang);
    if (total) {
        SG_CUDA(ctx, cudaMemcpyAsync(db->d_desc, h_desc, total * 32, cudaMemcpyHostToDevice, ctx->stream));
        SG_CUDA(ctx, cudaMemcpyAsync(db->d_angle, h_angle, total * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    SG_CUDA(ctx, cudaMemcpyAsync(db->d_offsets, db->offsets.data(), sizeof(long long) * ((size_t)n_sets + 1), cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors behind h_desc / h_angle may die with the caller's scope
    return SG_OK;
}

int sg_db_create(sg_ctx *ctx, const uint32_t *h_desc, const float *h_angle, const int64_t *h_offsets, int n_sets, sg_db **out) {
    return db_create(ctx, h_desc, h_angle, h_offsets, n_sets, out, cudaMemcpyHostToDevice);
}
int sg_db_create_device(sg_ctx *ctx, const uint32_t *d_desc, const float *d_angle, const int64_t *h_offsets, int n_sets, sg_db **out) {
    return db_create(ctx, d_desc, d_angle, h_offsets, n_sets, out, cudaMemcpyDeviceToDevice);
}
// A view: the descriptors stay where they are (e.g. the output of sg_extract_device); only the offsets are uploaded.
int sg_db_wrap_device(sg_ctx *ctx, const uint32_t *d_desc, const float *d_angle, const int64_t *h_offsets, int n_sets, sg_db **out) {
    cudaSetDevice(ctx->device);
    if (!d_desc || !d_angle || !h_offsets || n_sets < 1 || !out) return fail(ctx, SG_ERR_INVALID, "null / empty database");
    if (h_offsets[0] != 0) return fail(ctx, SG_ERR_INVALID, "offsets[0] must be 0");
    sg_db *db = new sg_db();
    db->owns_data = false;
    if (int r = db_finish(ctx, db, h_offsets, n_sets)) { sg_db_destroy(db); return r; }
    db->d_desc = const_cast<uint32_t *>(d_desc);
    db->d_angle = const_cast<float *>(d_angle);
    *out = db;
    return SG_OK;
}
void sg_db_destroy(sg_db *db) {
    if (!db) return;
    if (db->ctx) {
        cudaSetDevice(db->ctx->device);
        auto &v = db->ctx->dbs;
        v.erase(std::remove(v.begin(), v.end(), db), v.end());
    }
    auto release = [&](void *p) {      // stream-ordered, after everything queued on the context's stream so far
        if (!p) return;
        if (db->ctx) cudaFreeAsync(p, db->ctx->main_stream);
        else cudaFree(p);
    };
    if (db->owns_data) { release(db->d_desc); release(db->d_angle); }
    release(db->d_offsets);
    delete db;
}

int sg_match_pairs_device(sg_ctx *ctx, const sg_db *db, const int32_t *d_pairs, int n_pairs, const sg_match_params *mp,
                          int32_t *d_matches, int match_stride, uint32_t *d_n_matches) {
    cudaSetDevice(ctx->device);
    if (!db || !d_pairs || !mp || !d_n_matches) return fail(ctx, SG_ERR_INVALID, "null argument");
    return run_match(ctx, db, d_pairs, n_pairs, *mp, d_matches, match_stride, d_n_matches);
}

int sg_match_pairs(sg_ctx *ctx, const sg_db *db, const int32_t *h_pairs, int n_pairs, const sg_match_params *mp,
                   int32_t *h_matches, int match_stride, uint32_t *h_n_matches) {
    cudaSetDevice(ctx->device);
    if (!db || !h_pairs || !mp || !h_n_matches) return fail(ctx, SG_ERR_INVALID, "null argument");
    if (n_pairs <= 0) return SG_OK;
    for (int i = 0; i < 2 * n_pairs; ++i)
        if (h_pairs[i] < 0 || h_pairs[i] >= db->n_sets) return fail(ctx, SG_ERR_INVALID, "pair %d names set %d outside the database", i / 2, h_pairs[i]);
    if (h_matches && match_stride < db->max_set) return fail(ctx, SG_ERR_INVALID, "match_stride smaller than the largest set");
    if (int r = grow(ctx, (void **)&ctx->d_pairs, &ctx->pairs_cap, 2 * (size_t)n_pairs, sizeof(int))) return r;
    if (int r = grow(ctx, (void **)&ctx->d_nmatch, &ctx->nmatch_cap, (size_t)n_pairs, sizeof(uint32_t))) return r;
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->d_pairs, h_pairs, 2 * (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!h_matches) {
        if (int r = run_match(ctx, db, ctx->d_pairs, n_pairs, *mp, nullptr, 0, ctx->d_nmatch)) return r;
    } else {
        // Slabs of pairs through two device row buffers: the kernels of slab i+1 are queued before the copy-out of
        // slab i is issued on the second copy stream, so the copies (also the blocking ones into pageable memory) run
        // under the next slab's kernels.
        const int slab = std::min(n_pairs, std::min(256, match_chunk_pairs(db, false)));
        const size_t slab_rows = (size_t)slab * db->max_set;
        {
            size_t cap = ctx->matches_cap;
            const int r = grow(ctx, (void **)&ctx->d_matches, &cap, 2 * slab_rows, sizeof(int));
            ctx->matches_cap = cap;
            if (r) return r;
        }
        while ((int)ctx->pipe_ev.size() < 4) {
            cudaEvent_t e;
            SG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->pipe_ev.push_back(e);
        }
        if (match_stride > db->max_set)                  // -1 padding behind the largest set (the copies never touch it)
            for (int p = 0; p < n_pairs; ++p)
                for (int i = db->max_set; i < match_stride; ++i) h_matches[(size_t)p * match_stride + i] = -1;
        auto copy_out = [&](int k) -> int {              // slab k: rows of pairs [k * slab, ...)
            const int p0 = k * slab, np = std::min(slab, n_pairs - p0);
            SG_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->pipe_ev[k & 1], 0));
            SG_CUDA(ctx, cudaMemcpy2DAsync(h_matches + (size_t)p0 * match_stride, (size_t)match_stride * 4,
                                           ctx->d_matches + (k & 1) * slab_rows, (size_t)db->max_set * 4, (size_t)db->max_set * 4,
                                           np, cudaMemcpyDeviceToHost, ctx->s_out));
            SG_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[2 + (k & 1)], ctx->s_out));
            return SG_OK;
        };
        const int n_slabs = (n_pairs + slab - 1) / slab;
        for (int k = 0; k < n_slabs; ++k) {
            const int p0 = k * slab, np = std::min(slab, n_pairs - p0);
            if (k >= 2) SG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 + (k & 1)], 0));   // buffer free again
            if (int r = run_match(ctx, db, ctx->d_pairs + 2 * (size_t)p0, np, *mp, ctx->d_matches + (k & 1) * slab_rows, db->max_set,
                                  ctx->d_nmatch + p0, k == 0))
                return r;
            SG_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[k & 1], ctx->stream));
            if (k >= 1)
                if (int r = copy_out(k - 1)) return r;
        }
        if (int r = copy_out(n_slabs - 1)) return r;
        SG_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    }
    SG_CUDA(ctx, cudaMemcpyAsync(h_n_matches, ctx->d_nmatch, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(&ctx->rescans, ctx->d_rescans, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}

int sg_match_bruteforce(sg_ctx *ctx, const uint32_t *h_descA, const float *h_angA, int nA, const uint32_t *h_descB,
                        const float *h_angB, int nB, const sg_match_params *mp, int32_t *h_matches, uint32_t *n_matches) {
    cudaSetDevice(ctx->device);
    if (nA < 0 || nB < 0 || !mp || !n_matches || (nA && (!h_descA || !h_angA || !h_matches)) || (nB && (!h_descB || !h_angB)))
        return fail(ctx, SG_ERR_INVALID, "bad argument");
    *n_matches = 0;
    if (nA == 0) return SG_OK;
    if (nB == 0) { for (int i = 0; i < nA; ++i) h_matches[i] = -1; return SG_OK; }
    std::vector<uint32_t> desc(8 * ((size_t)nA + nB));
    std::vector<float> ang((size_t)nA + nB);
    memcpy(desc.data(), h_descA, 32 * (size_t)nA);
    memcpy(desc.data() + 8 * (size_t)nA, h_descB, 32 * (size_t)nB);
    memcpy(ang.data(), h_angA, 4 * (size_t)nA);
    memcpy(ang.data() + nA, h_angB, 4 * (size_t)nB);
    const int64_t offs[3] = {0, nA, (int64_t)nA + nB};
    sg_db tmp, *db = &tmp;
    if (int r = scratch_db(ctx, db, desc.data(), ang.data(), offs, 2)) return r;
    const int32_t pair[2] = {0, 1};
    std::vector<int32_t> m(db->max_set);
    const int r = sg_match_pairs(ctx, db, pair, 1, mp, m.data(), db->max_set, n_matches);
    if (!r) memcpy(h_matches, m.data(), 4 * (size_t)nA);
    return r;
}

// matchForLoopClosures with the reference's DBoW2 node buckets (keyframe_matcher.cpp:50-158).  Keypoints of
// different nodes never meet, so the sequential "already matched in kf2" state is per node: every node shared by
// the two keyframes becomes one small (A-part, B-part) pair of the batched brute-force matcher, and the angle
// histogram -- the one step that spans the nodes -- runs over the gathered matches on the host with the library's
// restated std::sort order (sg_angle_bin_order).
int sg_match_bow(sg_ctx *ctx, const uint32_t *h_descA, const float *h_angA, const int32_t *h_nodeA, const uint8_t *h_eligA,
                 int nA, const uint32_t *h_descB, const float *h_angB, const int32_t *h_nodeB, const uint8_t *h_eligB,
                 int nB, const sg_match_params *mp, int32_t *h_matches, uint32_t *n_matches) {
    cudaSetDevice(ctx->device);
    if (nA < 0 || nB < 0 || !mp || !n_matches || (nA && (!h_descA || !h_angA || !h_nodeA || !h_matches))
        || (nB && (!h_descB || !h_angB || !h_nodeB)))
        return fail(ctx, SG_ERR_INVALID, "bad argument");
    *n_matches = 0;
    for (int i = 0; i < nA; ++i) h_matches[i] = -1;
    if (nA == 0 || nB == 0) return SG_OK;
    // DBoW2::FeatureVector order: nodes ascending, features of a node in index order
    std::vector<std::pair<int, int>> fa, fb;   // (node, feature)
    for (int i = 0; i < nA; ++i) if (h_nodeA[i] >= 0 && (!h_eligA || h_eligA[i])) fa.emplace_back(h_nodeA[i], i);
    for (int i = 0; i < nB; ++i) if (h_nodeB[i] >= 0 && (!h_eligB || h_eligB[i])) fb.emplace_back(h_nodeB[i], i);
    std::stable_sort(fa.begin(), fa.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
    std::stable_sort(fb.begin(), fb.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
    std::vector<uint32_t> desc;
    std::vector<float> ang;
    std::vector<int64_t> offs{0};
    std::vector<int32_t> pairs, idxA, idxB;     // feature index of every database row
    std::vector<int> startA, startB;            // first row of every shared node inside idxA / idxB
    size_t ia = 0, ib = 0;
    while (ia < fa.size() && ib < fb.size()) {
        if (fa[ia].first < fb[ib].first) { ++ia; continue; }
        if (fb[ib].first < fa[ia].first) { ++ib; continue; }
        const int node = fa[ia].first;
        const int set = (int)offs.size() - 1;
        startA.push_back((int)idxA.size());
        for (; ia < fa.size() && fa[ia].first == node; ++ia) {
            const int i = fa[ia].second;
            desc.insert(desc.end(), h_descA + 8 * (size_t)i, h_descA + 8 * (size_t)i + 8);
            ang.push_back(h_angA[i]);
            idxA.push_back(i);
        }
        offs.push_back((int64_t)ang.size());
        startB.push_back((int)idxB.size());
        for (; ib < fb.size() && fb[ib].first == node; ++ib) {
            const int i = fb[ib].second;
            desc.insert(desc.end(), h_descB + 8 * (size_t)i, h_descB + 8 * (size_t)i + 8);
            ang.push_back(h_angB[i]);
            idxB.push_back(i);
        }
        offs.push_back((int64_t)ang.size());
        pairs.push_back(set);
        pairs.push_back(set + 1);
    }
    const int n_nodes = (int)pairs.size() / 2;
    if (n_nodes == 0) return SG_OK;
    sg_db tmp, *db = &tmp;
    if (int r = scratch_db(ctx, db, desc.data(), ang.data(), offs.data(), (int)offs.size() - 1)) return r;
    sg_match_params q = *mp;
    q.check_orientation = 0;                    // the histogram spans all nodes: applied below
    const int stride = std::max(db->max_set, 1);
    std::vector<int32_t> rows((size_t)n_nodes * stride);
    std::vector<uint32_t> counts(n_nodes);
    if (int r = sg_match_pairs(ctx, db, pairs.data(), n_nodes, &q, rows.data(), stride, counts.data())) return r;
    uint32_t num = 0;
    for (int k = 0; k < n_nodes; ++k) {
        const int na = (int)(offs[2 * k + 1] - offs[2 * k]);
        for (int j = 0; j < na; ++j) {
            const int m = rows[(size_t)k * stride + j];
            if (m >= 0) { h_matches[idxA[startA[k] + j]] = idxB[startB[k] + m]; ++num; }
        }
    }
    if (mp->check_orientation) {                // match_angle_checker.h:72-134
        uint32_t sizes[30] = {0}, order[30];
        std::vector<int> bin(nA, -1);
        for (int i = 0; i < nA; ++i)
            if (h_matches[i] >= 0) { bin[i] = sg_angle_bin(h_angA[i] - h_angB[h_matches[i]]); ++sizes[bin[i]]; }
        sg_angle_bin_order(sizes, order);
        for (int i = 0; i < nA; ++i)
            if (bin[i] >= 0 && (uint32_t)bin[i] != order[0] && (uint32_t)bin[i] != order[1] && (uint32_t)bin[i] != order[2]) {
                h_matches[i] = -1;
                --num;
            }
    }
    *n_matches = num;
    return SG_OK;
}

// matchForTriangulationDBoW (keyframe_matcher.cpp:160-293).  The O(nA * nB) Hamming work runs on the GPU (per node and
// kf1 feature the candidate kf2 features within thr, ascending distance, ties by descending index); the acceptance walks
// those short lists on the host: uniqueness in kf2 and the fp64 epipolar test (:23-44: acos / sqrt from the host libm, as
// the reference evaluates it).  A list that is truncated and exhausted is completed exactly by a second GPU pass.
namespace {
inline double sum3(double p0, double p1, double p2) { return p0 + (p1 + p2); }   // Eigen's fixed-size redux order
bool epipolar_ok(const double *b1, const double *b2, const double *E, float scale, float residual_deg_thr) {
    const double e[3] = {sum3(E[0] * b2[0], E[1] * b2[1], E[2] * b2[2]), sum3(E[3] * b2[0], E[4] * b2[1], E[5] * b2[2]),
                         sum3(E[6] * b2[0], E[7] * b2[1], E[8] * b2[2])};
    const double cos_residual = sum3(e[0] * b1[0], e[1] * b1[1], e[2] * b1[2]) / std::sqrt(sum3(e[0] * e[0], e[1] * e[1], e[2] * e[2]));
    const double residual_rad = M_PI / 2.0 - std::abs(std::acos(cos_residual));
    const double residual_rad_thr = residual_deg_thr * M_PI / 180.0;
    return residual_rad < residual_rad_thr * scale;
}
}  // namespace

int sg_match_triangulation(sg_ctx *ctx, const uint32_t *h_descA, const float *h_angA, const int32_t *h_octA, const double *h_bearA,
                           const int32_t *h_nodeA, const uint8_t *h_eligA, int nA, const uint32_t *h_descB, const float *h_angB,
                           const double *h_bearB, const int32_t *h_nodeB, const uint8_t *h_eligB, int nB,
                           const sg_triangulation_params *tp, int32_t *h_matches, uint32_t *n_matches) {
    cudaSetDevice(ctx->device);
    if (nA < 0 || nB < 0 || !tp || !n_matches || !tp->scale_factors
        || (nA && (!h_descA || !h_angA || !h_octA || !h_bearA || !h_nodeA || !h_matches))
        || (nB && (!h_descB || !h_angB || !h_bearB || !h_nodeB)))
        return fail(ctx, SG_ERR_INVALID, "bad argument");
    if (tp->thr > 256) return fail(ctx, SG_ERR_INVALID, "thr must be <= 256");
    *n_matches = 0;
    for (int i = 0; i < nA; ++i) {
        h_matches[i] = -1;
        if (h_octA[i] < 0 || h_octA[i] >= tp->n_levels) return fail(ctx, SG_ERR_INVALID, "octave of feature %d outside the scale factors", i);
    }
    if (nA == 0 || nB == 0) return SG_OK;
    std::vector<std::pair<int, int>> fa, fb;   // (node, feature), DBoW2::FeatureVector order
    for (int i = 0; i < nA; ++i) if (h_nodeA[i] >= 0 && (!h_eligA || h_eligA[i])) fa.emplace_back(h_nodeA[i], i);
    for (int i = 0; i < nB; ++i) if (h_nodeB[i] >= 0 && (!h_eligB || h_eligB[i])) fb.emplace_back(h_nodeB[i], i);
    std::stable_sort(fa.begin(), fa.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
    std::stable_sort(fb.begin(), fb.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
    std::vector<uint32_t> desc;
    std::vector<float> ang;
    std::vector<int64_t> offs{0};
    std::vector<int32_t> pairs, idxA, idxB;
    std::vector<int> startA, startB;
    size_t ia = 0, ib = 0;
    while (ia < fa.size() && ib < fb.size()) {
        if (fa[ia].first < fb[ib].first) { ++ia; continue; }
        if (fb[ib].first < fa[ia].first) { ++ib; continue; }
        const int node = fa[ia].first, set = (int)offs.size() - 1;
        startA.push_back((int)idxA.size());
        for (; ia < fa.size() && fa[ia].first == node; ++ia) {
            const int i = fa[ia].second;
            desc.insert(desc.end(), h_descA + 8 * (size_t)i, h_descA + 8 * (size_t)i + 8);
            ang.push_back(h_angA[i]);
            idxA.push_back(i);
        }
        offs.push_back((int64_t)ang.size());
        startB.push_back((int)idxB.size());
        for (; ib < fb.size() && fb[ib].first == node; ++ib) {
            const int i = fb[ib].second;
            desc.insert(desc.end(), h_descB + 8 * (size_t)i, h_descB + 8 * (size_t)i + 8);
            ang.push_back(h_angB[i]);
            idxB.push_back(i);
        }
        offs.push_back((int64_t)ang.size());
        pairs.push_back(set);
        pairs.push_back(set + 1);
    }
    const int n_nodes = (int)pairs.size() / 2;
    if (n_nodes == 0) return SG_OK;
    sg_db tmp, *db = &tmp;
    if (int r = scratch_db(ctx, db, desc.data(), ang.data(), offs.data(), (int)offs.size() - 1)) return r;
    if (int r = grow(ctx, (void **)&ctx->d_pairs, &ctx->pairs_cap, pairs.size(), sizeof(int))) return r;
    SG_CUDA(ctx, cudaMemcpyAsync(ctx->d_pairs, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *d_topk = nullptr, *d_nseen = nullptr;
    int stride = 0;
    if (int r = run_topk_lists(ctx, db, ctx->d_pairs, n_nodes, tp->thr, &d_topk, &d_nseen, &stride)) return r;
    std::vector<uint32_t> topk((size_t)n_nodes * stride * 4), nseen((size_t)n_nodes * stride);
    SG_CUDA(ctx, cudaMemcpyAsync(topk.data(), d_topk, topk.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaMemcpyAsync(nseen.data(), d_nseen, nseen.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    std::vector<char> taken(nB, 0);
    uint32_t num = 0;
    unsigned long long rescans = 0;
    std::vector<uint32_t> full;     // complete candidate list of a row whose top-4 list ran dry
    for (int k = 0; k < n_nodes; ++k) {
        const int na = (int)(offs[2 * k + 1] - offs[2 * k]), nb = (int)(offs[2 * k + 2] - offs[2 * k + 1]);
        for (int j = 0; j < na; ++j) {
            const int i1 = idxA[startA[k] + j];
            const size_t r = (size_t)k * stride + j;
            const uint32_t ns = nseen[r];
            if (ns == 0) continue;
            const float scale = tp->scale_factors[h_octA[i1]];
            int best = -1;
            // the reference keeps the LAST candidate with d <= best-so-far that passes: the smallest distance wins,
            // among equal distances the largest index -- exactly the order of the keys
            auto try_key = [&](uint32_t key) {
                const int m = (int)(0xffffu - (key & 0xffffu));
                const int i2 = idxB[startB[k] + m];
                if (taken[i2]) return false;
                if (!epipolar_ok(h_bearA + 3 * (size_t)i1, h_bearB + 3 * (size_t)i2, tp->E, scale, tp->residual_deg_thr)) return false;
                best = i2;
                return true;
            };
            bool done = false;
            for (int e = 0; e < 4 && !done; ++e) {
                const uint32_t key = topk[4 * r + e];
                if (key == 0xffffffffu) break;
                done = try_key(key);
            }
            if (!done && ns > 4) {
                // truncated and exhausted: exact rescan of this row on the GPU (all candidates within thr)
                ++rescans;
                const int row[2] = {(int)(offs[2 * k] + j), 2 * k + 1};
                if (int rr = grow(ctx, &ctx->d_tmp, &ctx->tmp_bytes, 8 + 4 * ((size_t)nb + 1), 1)) return rr;
                int *d_row = (int *)ctx->d_tmp;
                uint32_t *d_out = (uint32_t *)((uint8_t *)ctx->d_tmp + 8), *d_cnt = d_out + nb;
                SG_CUDA(ctx, cudaMemcpyAsync(d_row, row, 8, cudaMemcpyHostToDevice, ctx->stream));
                if (int rr = run_row_scan(ctx, db, d_row, 1, tp->thr, nb, d_out, d_cnt)) return rr;
                full.resize((size_t)nb + 1);
                SG_CUDA(ctx, cudaMemcpyAsync(full.data(), d_out, 4 * ((size_t)nb + 1), cudaMemcpyDeviceToHost, ctx->stream));
                SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                const uint32_t cnt = full[nb];
                for (uint32_t c = 0; c < cnt; ++c) full[c] = (full[c] & 0xffff0000u) | (0xffffu - (full[c] & 0xffffu));
                std::sort(full.begin(), full.begin() + cnt);
                for (uint32_t c = 0; c < cnt && !done; ++c) done = try_key(full[c]);
            }
            if (best < 0) continue;
            taken[best] = 1;
            h_matches[i1] = best;
            ++num;
        }
    }
    ctx->rescans = rescans;
    if (tp->check_orientation) {                // match_angle_checker.h:72-134
        uint32_t sizes[30] = {0}, order[30];
        std::vector<int> bin(nA, -1);
        for (int i = 0; i < nA; ++i)
            if (h_matches[i] >= 0) { bin[i] = sg_angle_bin(h_angA[i] - h_angB[h_matches[i]]); ++sizes[bin[i]]; }
        sg_angle_bin_order(sizes, order);
        for (int i = 0; i < nA; ++i)
            if (bin[i] >= 0 && (uint32_t)bin[i] != order[0] && (uint32_t)bin[i] != order[1] && (uint32_t)bin[i] != order[2]) {
                h_matches[i] = -1;
                --num;
            }
    }
    *n_matches = num;
    return SG_OK;
}

unsigned long long sg_match_rescans(const sg_ctx *ctx) { return ctx->rescans; }

// ---- helpers --------------------------------------------------------------------------------------------
int sg_malloc(sg_ctx *ctx, size_t bytes, void **d_ptr) {
    cudaSetDevice(ctx->device);
    SG_CUDA(ctx, cudaMalloc(d_ptr, bytes));
    return SG_OK;
}
int sg_free(sg_ctx *ctx, void *d_ptr) {
    cudaSetDevice(ctx->device);
    SG_CUDA(ctx, cudaFree(d_ptr));
    return SG_OK;
}
int sg_memcpy_h2d(sg_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    cudaSetDevice(ctx->device);
    SG_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}
int sg_memcpy_d2h(sg_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    cudaSetDevice(ctx->device);
    SG_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SG_OK;
}
int sg_host_alloc_pinned(size_t bytes, void **h_ptr) { return cudaMallocHost(h_ptr, bytes) == cudaSuccess ? SG_OK : SG_ERR_CUDA; }
int sg_host_free_pinned(void *h_ptr) { return cudaFreeHost(h_ptr) == cudaSuccess ? SG_OK : SG_ERR_CUDA; }

int sg_timer_start(sg_ctx *ctx) {
    SG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return SG_OK;
}
int sg_timer_stop(sg_ctx *ctx, float *ms) {
    SG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    SG_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    SG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return SG_OK;
}
int sg_flush_l2(sg_ctx *ctx) {
    cudaSetDevice(ctx->device);
    const size_t bytes = (size_t)256 << 20;   // 2x the 126 MB L2
    if (!ctx->d_flush) {
        SG_CUDA(ctx, cudaMalloc(&ctx->d_flush, bytes));
        ctx->flush_bytes = bytes;
    }
    SG_CUDA(ctx, cudaMemsetAsync(ctx->d_flush, 0x5a, ctx->flush_bytes, ctx->stream));
    return SG_OK;
}
int sg_set_profiling(sg_ctx *ctx, int on) {
    ctx->profiling = on != 0;
    ctx->prof_call = -1;
    memset(ctx->stage_mark, 0, sizeof ctx->stage_mark);
    return SG_OK;
}
int sg_get_stage_ms(sg_ctx *ctx, float *ms6, int *n_calls) {
    cudaSetDevice(ctx->device);
    if (!ms6) return fail(ctx, SG_ERR_INVALID, "null argument");
    SG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int from[6] = {EV_PYR0, EV_PYR1, EV_FAST1, EV_DIST1, EV_MATCH0, EV_TOPK1};
    const int to[6] = {EV_PYR1, EV_FAST1, EV_DIST1, EV_DESC1, EV_TOPK1, EV_RESOLVE1};
    const long calls = std::min<long>(ctx->prof_call + 1, sg_ctx::PROF_SLOTS);
    for (int i = 0; i < 6; ++i) {
        double sum = 0;
        int n = 0;
        for (long c = 0; c < calls; ++c) {
            if (!ctx->stage_mark[c][from[i]] || !ctx->stage_mark[c][to[i]]) continue;
            float ms = 0;
            SG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_stage[c][from[i]], ctx->ev_stage[c][to[i]]));
            sum += ms;
            ++n;
        }
        ms6[i] = n ? (float)(sum / n) : -1.f;
    }
    if (n_calls) *n_calls = (int)calls;
    return SG_OK;
}
int sg_microbench_popc(sg_ctx *ctx, double *popc_per_s, float *ms) {
    cudaSetDevice(ctx->device);
    if (!popc_per_s || !ms) return fail(ctx, SG_ERR_INVALID, "null argument");
    return run_popc_bench(ctx, popc_per_s, ms);
}

}  // extern "C"
