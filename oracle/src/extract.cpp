// ORACLE (test infrastructure only): OrbExtractor::detectAndExtract for one frame and threaded
// CPU-baseline drivers.  Follows orb_extractor.cpp:73-164 (tracker branch :89-124, detection :128,
// orientation :132-134, descriptors + output assembly :139-163) and feature_detector.cpp:103-133.
#include "common.h"
#include <chrono>
#include <malloc.h>
#include <thread>
#include <atomic>

using namespace orc;

extern "C" unsigned orc_match_bruteforce(const uint32_t *, const float *, int, const uint32_t *, const float *, int,
                                         float, unsigned, int, int, int *);

extern "C" int orc_extract(const orc_params *p, const uint8_t *img, int stride,
                           const float *track_xy, const int *track_ids, int n_tracks, int track_level,
                           float *x, float *y, float *angle, int *octave, uint32_t *desc, int *track_id,
                           int *lvl_x, int *lvl_y, int cap, int *level_counts) {
    const Geometry g = make_geometry(*p);
    // per-thread planes, kept across calls like the reference's cached pyramid (orb_extractor.cpp:201-203): allocating
    // and faulting in 2 x 0.95 MB per frame serialises the threads of the CPU baseline on the kernel's mmap lock
    thread_local std::vector<uint8_t> pyr, blur;
    if (pyr.size() < g.off[g.levels]) { pyr.resize(g.off[g.levels]); blur.resize(g.off[g.levels]); }
    orc_pyramid(p, img, stride, pyr.data(), blur.data());
    int n = 0;
    auto emit = [&](float fx, float fy, float a, int oct, const uint32_t *d, int tid, int lx, int ly) {
        if (n < cap) {
            x[n] = fx; y[n] = fy; angle[n] = a; octave[n] = oct;
            memcpy(desc + 8 * (size_t)n, d, 32);
            if (track_id) track_id[n] = tid;
            if (lvl_x) lvl_x[n] = lx;
            if (lvl_y) lvl_y[n] = ly;
        }
        ++n;
    };
    // tracker keypoints (orb_extractor.cpp:89-124); camera.isValidPixel == true
    for (int t = 0; t < n_tracks; ++t) {
        const float px = track_xy[2 * t], py = track_xy[2 * t + 1];
        const int level = track_level;
        const float scale = g.scale[level];
        const uint8_t *lev = pyr.data() + g.off[level];
        const int ix = cv_round(px / scale), iy = cv_round(py / scale);
        const int margin = PATCH_RADIUS;
        if (ix >= margin && iy >= margin && ix < g.w[level] - margin && iy < g.h[level] - margin) {
            const float a = orc_ic_angle(lev, g.w[level], ix, iy);
            uint32_t d[8];
            orc_descriptor(blur.data() + g.off[level], g.w[level], ix, iy, a, d);
            emit(px, py, a, level, d, track_ids ? track_ids[t] : t, ix, iy);
        }
    }
    // detected keypoints, level by level
    for (int l = 0; l < g.levels; ++l) {
        const uint8_t *lev = pyr.data() + g.off[l];
        const std::vector<Kp> kps = detect_level(lev, g.w[l], g.h[l], g.w[l], g.budget[l],
                                                 p->ini_fast_thr, p->min_fast_thr, nullptr);
        int cnt = 0;
        for (const Kp &k : kps) {
            const float kx = (float)k.x, ky = (float)k.y;
            // feature_detector.cpp:117-123 border filter (std::round, half away from zero)
            const int rx = (int)std::round(kx), ry = (int)std::round(ky);
            if (rx < PATCH_RADIUS || ry < PATCH_RADIUS || rx >= g.w[l] - PATCH_RADIUS || ry >= g.h[l] - PATCH_RADIUS)
                continue;
            const int ix = cv_round(kx), iy = cv_round(ky);  // orb_extractor.cpp:241,280
            const float a = orc_ic_angle(lev, g.w[l], ix, iy);
            uint32_t d[8];
            orc_descriptor(blur.data() + g.off[l], g.w[l], ix, iy, a, d);
            emit(kx * g.scale[l], ky * g.scale[l], a, l, d, -1, ix, iy);  // :153-162
            ++cnt;
        }
        if (level_counts) level_counts[l] = cnt;
    }
    return n;
}

// glibc returns every freed block above 128 KB to the kernel (munmap) and maps the next one afresh: the per-call
// temporaries of the pyramid / FAST code (0.3 - 2 MB) then cost a page-fault storm per frame and the threads of the
// baseline serialise on the process's mmap lock (measured: 8 threads 4.3x slower).  Keep such blocks in the heap.
extern "C" void orc_tune_malloc(void) {
    mallopt(M_MMAP_THRESHOLD, 32 << 20);
    mallopt(M_TRIM_THRESHOLD, 512 << 20);
    mallopt(M_TOP_PAD, 64 << 20);
}

template <class F>
static double run_threads(int n_items, int threads, F fn) {
    orc_tune_malloc();
    threads = std::max(1, threads);
    std::atomic<int> next{0};
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (int i; (i = next.fetch_add(1)) < n_items;) fn(i);
        });
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

extern "C" double orc_bench_extract(const orc_params *p, const uint8_t *imgs, int n_frames, int threads, long *total_kp) {
    std::atomic<long> total{0};
    const size_t fsz = (size_t)p->width * p->height;
    const int cap = p->max_keypoints + 64;
    const double s = run_threads(n_frames, threads, [&](int i) {
        std::vector<float> x(cap), y(cap), a(cap);
        std::vector<int> o(cap);
        std::vector<uint32_t> d(8 * (size_t)cap);
        total += orc_extract(p, imgs + fsz * i, p->width, nullptr, nullptr, 0, 0,
                             x.data(), y.data(), a.data(), o.data(), d.data(), nullptr, nullptr, nullptr, cap, nullptr);
    });
    if (total_kp) *total_kp = total;
    return s;
}

extern "C" double orc_bench_match(const uint32_t *desc, const float *ang, int n_sets, int n_per_set,
                                  const int *pairs, int n_pairs, float ratio, unsigned thr, int threads,
                                  long *total_matches) {
    (void)n_sets;
    std::atomic<long> total{0};
    const double s = run_threads(n_pairs, threads, [&](int i) {
        std::vector<int> m(n_per_set);
        const int a = pairs[2 * i], b = pairs[2 * i + 1];
        total += orc_match_bruteforce(desc + (size_t)a * n_per_set * 8, ang + (size_t)a * n_per_set, n_per_set,
                                      desc + (size_t)b * n_per_set * 8, ang + (size_t)b * n_per_set, n_per_set,
                                      ratio, thr, 1, 0, m.data());
    });
    if (total_matches) *total_matches = total;
    return s;
}
