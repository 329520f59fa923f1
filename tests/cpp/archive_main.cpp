// Driver of the map-archive reader / writer (slam-module_b200/host/map_archive.cpp) for tests/test_map_archive.py:
//   archive_main roundtrip <in> <out> [fixed_header]   load, print a summary, save again (no GPU needed)
//   archive_main match <in> [fixed_header]             load, rebuild the descriptor database on the GPU and match the
//                                                      first two keyframes with sg_match_pairs (loadMapDB's rebuild hook)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/slamgpu.h"
#include "../../slam-module_b200/host/slam_frontend.hpp"

using namespace slam;

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s roundtrip <in> <out> [fixed_header] | match <in> [fixed_header]\n", argv[0]); return 2; }
    const std::string cmd = argv[1];
    MapArchiveOptions opt;
    opt.fixedSizeHeader = (cmd == "roundtrip" ? argc > 4 : argc > 3);
    MapDB db;
    std::string err;
    if (!loadMapArchive(argv[2], db, &err, opt)) { std::printf("error: %s\n", err.c_str()); return 1; }
    size_t nkp = 0;
    for (const auto &e : db.keyframes) nkp += e.second && e.second->shared ? e.second->shared->keyPoints.size() : 0;
    std::printf("keyframes %zu keypoints %zu mappoints %zu tracks %zu edges %zu nextMp %d lastKf %d t0 %.3f\n", db.keyframes.size(), nkp,
                db.mapPoints.size(), db.trackIdToMapPoint.size(), db.loopClosureEdges.size(), db.nextMp, db.lastKfId.v, db.firstKfTimestamp);
    for (const auto &e : db.keyframes) {
        const Keyframe &kf = *e.second;
        std::printf("kf %d prev %d next %d kps %zu camera '%s' t %.3f pose03 %.3f full %d", kf.id.v, kf.previousKfId.v, kf.nextKfId.v,
                    kf.shared->keyPoints.size(), kf.shared->cameraModel.c_str(), kf.t, kf.poseCW(0, 3), (int)kf.hasFullFeatures);
        if (!kf.shared->keyPoints.empty()) {
            const KeyPoint &kp = kf.shared->keyPoints[0];
            std::printf(" kp0 %.2f %.2f %.2f %d %.3f %08x", kp.pt.x, kp.pt.y, kp.angle, kp.octave, kp.bearing(2), kp.descriptor[7]);
        }
        std::printf("\n");
    }
    for (const auto &e : db.mapPoints)
        std::printf("mp %d status %d obs %zu pos %.3f %.3f %.3f ref %d\n", e.second.id.v, (int)e.second.status, e.second.observations.size(),
                    e.second.position(0), e.second.position(1), e.second.position(2), e.second.referenceKeyframe.v);
    if (cmd == "roundtrip") {
        if (!saveMapArchive(argv[3], db, &err, opt)) { std::printf("error: %s\n", err.c_str()); return 1; }
        return 0;
    }
    // ---- rebuild hook on the GPU -------------------------------------------------------------------------------------
    sg_params p{};
    p.width = 640; p.height = 480; p.levels = 8; p.scale_factor = 1.2f; p.max_keypoints = 1000; p.ini_fast_thr = 20; p.min_fast_thr = 7;
    p.max_frames = 1;
    sg_ctx *ctx = nullptr;
    if (sg_create(0, &p, &ctx) != SG_OK) { std::printf("error: %s\n", sg_last_error(nullptr)); return 3; }
    std::vector<KfId> ids;
    sg_db *sdb = buildDescriptorDatabase(db, ctx, ids);
    if (!sdb || ids.size() < 2) { std::printf("error: fewer than two keyframes\n"); return 4; }
    const std::int32_t pair[2] = {0, 1};
    const size_t n0 = db.keyframes.at(ids[0])->shared->keyPoints.size();
    std::vector<std::int32_t> m(n0 ? n0 : 1, -1);
    std::uint32_t n = 0;
    sg_match_params mp{};
    mp.ratio = 0.8f; mp.thr = 50; mp.check_orientation = 1;
    if (sg_match_pairs(ctx, sdb, pair, 1, &mp, m.data(), (int)m.size(), &n) != SG_OK) { std::printf("error: %s\n", sg_last_error(ctx)); return 5; }
    std::printf("matched %u of %zu between kf %d and kf %d:", n, n0, ids[0].v, ids[1].v);
    for (size_t i = 0; i < n0; ++i) std::printf(" %d", m[i]);
    std::printf("\n");
    std::vector<size_t> around;
    db.keyframes.at(ids[0])->getFeaturesAround(Vector2f(0.f, 0.f), 1e9f, around);
    std::printf("feature search rebuilt: %zu\n", around.size());
    sg_db_destroy(sdb);
    sg_destroy(ctx);
    return 0;
}
