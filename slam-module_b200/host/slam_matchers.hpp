// The descriptor matchers of keyframe_matcher.hpp:53-91 and the batched MapPoint::updateDescriptor
// (map_point.cpp:75-116) on top of the C ABI of libslamgpu.so, as templates over the CALLER'S Keyframe / MapPoint /
// MapDB / StaticSettings types.
//
// Why templates: the reference's data model (keyframe.hpp, map_point.hpp, mapdb.hpp) is not part of the hot path and is
// not rebuilt here.  These functions touch it only through the members the reference's own matchers touch
// (kf.reproject, kf.cameraCenter, kf.mapPoints, kf.shared->keyPoints, mp.position, mp.predictScaleLevel,
// mp.addObservation, mp.replaceWith, mapDB.mapPoints ...), so the same source instantiates against
//   * the stand-in types of slam_frontend.hpp (explicit instantiations in libslam_frontend.so), and
//   * the reference's real classes (tests/cpp/ref_adapter_main.cpp does that, next to the reference's own functions).
// Each function keeps the reference's three phases -- projection geometry on the host, the candidate loops
// (FeatureSearch radius query + Hamming best / second best) in ONE library call for all map points, bookkeeping on
// the host in the reference's order -- and therefore its exact results; see the notes on order dependence below.
//
// `Types` names the caller's linear-algebra types: Vector2f, Vector2d, Vector3d, Matrix3d, Matrix4d (Eigen's in the
// reference) and `static Matrix3d createE21(R1, t1, R2, t2)` (openvslam::solve::essential_solver::create_E_21).
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <utility>
#include <vector>

#include "../../include/slamgpu.h"

namespace slam {
namespace cuda_matchers {

inline void check(sg_ctx *ctx, int rc, const char *what) {
    if (rc == SG_OK) return;
    // the reference has no error codes on this path (asserts only): fail loudly
    std::fprintf(stderr, "slam-b200: %s failed (%d): %s\n", what, rc, sg_last_error(ctx));
    std::abort();
}

// Keypoints of a keyframe as the flat arrays the library takes; `order` = FeatureSearch's Y-sorted index
// (feature_search.cpp:22-31), produced by the library with the same std::sort call.
struct KeypointArrays {
    std::vector<float> x, y;
    std::vector<std::int32_t> octave, order;
    std::vector<std::uint32_t> desc;
    int n = 0;
};
template <class Keyframe> void gatherKeypoints(const Keyframe &kf, KeypointArrays &a) {
    const auto &kps = kf.shared->keyPoints;
    a.n = (int)kps.size();
    a.x.resize(a.n); a.y.resize(a.n); a.octave.resize(a.n); a.order.resize(a.n); a.desc.resize(8 * (size_t)a.n);
    for (int i = 0; i < a.n; ++i) {
        a.x[i] = kps[i].pt.x; a.y[i] = kps[i].pt.y; a.octave[i] = kps[i].octave;
        std::memcpy(&a.desc[8 * (size_t)i], kps[i].descriptor.data(), 32);
    }
    if (a.n) sg_feature_index(a.x.data(), a.y.data(), a.n, a.order.data());
}

struct QueryArrays {
    std::vector<float> x, y, r;
    std::vector<std::int32_t> level;
    std::vector<std::uint32_t> desc;
    void push(float qx, float qy, float qr, int lvl, const std::uint32_t *d) {
        x.push_back(qx); y.push_back(qy); r.push_back(qr); level.push_back(lvl);
        desc.insert(desc.end(), d, d + 8);
    }
    int size() const { return (int)x.size(); }
};

// ---- searchByProjection (keyframe_matcher.hpp:59-66, keyframe_matcher.cpp:295-414) ------------------------------
// Order dependence: a keypoint matched by an earlier map point is skipped by later ones (:358).  The library's mode 1
// resolves the queries in order with exactly that rule, so phase 3 only has to record the accepted matches.
// (The ViewerDataPublisher argument of the reference is a debug hook and is not taken.)
template <class Types, class Keyframe, class MpIdT, class MapDB, class Settings>
int searchByProjection(Keyframe &kf, const std::vector<MpIdT> &mps, MapDB &mapDB, const float threshold,
                       const Settings &settings, sg_ctx *ctx) {
    using KpIdT = typename std::decay<decltype(mapDB.mapPoints.begin()->second.observations.begin()->second)>::type;
    QueryArrays q;
    std::vector<MpIdT> live;
    const std::size_t refLevel = settings.scaleFactors.size() / 2;
    for (const MpIdT mpId : mps) {
        auto &mp = mapDB.mapPoints.at(mpId);
        typename Types::Vector2f pix;
        if (!kf.reproject(mp.position, pix)) continue;                                   // :314-318
        const typename Types::Vector3d toKf = kf.cameraCenter() - mp.position;
        const auto toKfF = toKf.template cast<float>();
        const float dist = toKfF.norm();
        if (dist < mp.minViewingDistance || mp.maxViewingDistance < dist) continue;      // :322-325
        const float cosView = toKfF.normalized().dot(mp.norm);
        if (cosView < 0.5f) continue;                                                    // :327-329
        const int level = mp.predictScaleLevel(dist, settings);
        const float shrink = cosView > 0.998f ? 2.5f / 4.0f : 1.0f;                      // :333-336
        const float radius = shrink * threshold * settings.scaleFactors.at((std::size_t)level) / settings.scaleFactors.at(refLevel);
        q.push(pix(0), pix(1), radius, level, mp.descriptor.data());
        live.push_back(mpId);
    }
    if (live.empty()) return 0;
    KeypointArrays k;
    gatherKeypoints(kf, k);
    std::vector<std::uint8_t> taken((std::size_t)std::max(k.n, 1), 0);
    for (int i = 0; i < k.n; ++i)                                                         // :358
        taken[i] = kf.mapPoints[i].v != -1 && mapDB.mapPoints.at(kf.mapPoints.at(i)).observations.size() > 0;
    std::vector<std::int32_t> idx(live.size());
    std::vector<std::uint32_t> dist(live.size());
    check(ctx, sg_search_candidates(ctx, k.x.data(), k.y.data(), k.octave.data(), k.desc.data(), k.n, k.order.data(), taken.data(),
                                    q.x.data(), q.y.data(), q.r.data(), q.desc.data(), nullptr, q.size(), 1, 100u, idx.data(),
                                    dist.data(), nullptr), "sg_search_candidates");
    int matchCount = 0;
    for (std::size_t i = 0; i < live.size(); ++i) {
        if (idx[i] < 0) continue;
        auto &mp = mapDB.mapPoints.at(live[i]);
        kf.addObservation(mp.id, KpIdT(idx[i]));                                          // :388-390
        mp.addObservation(kf.id, KpIdT(idx[i]));
        ++matchCount;
    }
    return matchCount;
}

// ---- replaceDuplication (keyframe_matcher.hpp:72-78, keyframe_matcher.cpp:416-529) -----------------------------
// The candidate loop (:482-494) keeps the best keypoint only and never consumes one, so the best keypoint of a map
// point does not depend on the fusing done for earlier map points: it is computed for every map point in one
// library call, then the reference's sequential walk (:424-431 dynamic skips, :500-524 fusing) runs unchanged.
template <class Types, class Keyframe, class Container, class MapDB, class Settings>
unsigned int replaceDuplication(Keyframe &kf, const Container &mapPoints, const float margin, MapDB &mapDB,
                                const Settings &settings, sg_ctx *ctx) {
    using MpIdT = typename std::decay<decltype(*mapPoints.begin())>::type;
    using KpIdT = typename std::decay<decltype(mapDB.mapPoints.begin()->second.observations.begin()->second)>::type;
    using Status = typename std::decay<decltype(mapDB.mapPoints.begin()->second.status)>::type;
    constexpr float SQRT_CHI2_INV2D = 2.4477;                                              // keyframe_matcher.cpp:17
    QueryArrays q;
    std::vector<int> queryOf;                  // per list position: query index or -1 (skipped by a static test)
    const float baseScale = settings.scaleFactors[settings.scaleFactors.size() / 2];
    for (const MpIdT &mpId : mapPoints) {
        queryOf.push_back(-1);
        if (mpId.v == -1 || !mapDB.mapPoints.count(mpId)) continue;
        const auto &mp = mapDB.mapPoints.at(mpId);
        if (mp.status == Status::BAD || mp.status == Status::NOT_TRIANGULATED) continue;   // :434-436
        typename Types::Vector2f pix;
        if (!kf.reproject(mp.position, pix)) continue;                                     // :439-443
        const typename Types::Vector3d toKf = kf.cameraCenter() - mp.position;
        const auto toKfF = toKf.template cast<float>();
        const float dist = toKfF.norm();
        if (dist < mp.minViewingDistance || mp.maxViewingDistance < dist) continue;        // :451-453
        if (mp.norm.isZero(0)) continue;                                                   // :456-458
        if (toKfF.normalized().dot(mp.norm) < 0.5) continue;                               // :460-462
        const int level = mp.predictScaleLevel(dist, settings);
        const float radius = margin * settings.scaleFactors[(std::size_t)level] / baseScale * SQRT_CHI2_INV2D;   // :467
        queryOf.back() = q.size();
        q.push(pix(0), pix(1), radius, level, mp.descriptor.data());
    }
    std::vector<std::int32_t> best((std::size_t)std::max(q.size(), 1), -1);
    if (q.size() > 0) {
        KeypointArrays k;
        gatherKeypoints(kf, k);
        std::vector<std::uint32_t> dist(best.size());
        check(ctx, sg_search_candidates(ctx, k.x.data(), k.y.data(), k.octave.data(), k.desc.data(), k.n, k.order.data(), nullptr,
                                        q.x.data(), q.y.data(), q.r.data(), q.desc.data(), nullptr, q.size(), 0, 50u, best.data(),
                                        dist.data(), nullptr), "sg_search_candidates");
    }
    std::set<MpIdT> erased;
    unsigned int fusedCount = 0;
    std::size_t pos = 0;
    for (const MpIdT &mpId : mapPoints) {
        const int qi = queryOf[pos++];
        if (mpId.v == -1 || erased.count(mpId)) continue;                                  // :424-426
        auto &mp = mapDB.mapPoints.at(mpId);
        if (mp.observations.count(kf.id)) continue;                                        // :429-431
        if (qi < 0 || best[(std::size_t)qi] < 0) continue;   // static skips / no keypoint within HAMMING_DIST_THR_LOW (:497-499)
        const KpIdT bestKp(best[(std::size_t)qi]);
        const MpIdT matchedId = kf.mapPoints[bestKp.v];
        if (matchedId.v == -1) {                                                           // :502-505
            mp.addObservation(kf.id, bestKp);
            kf.addObservation(mp.id, bestKp);
        } else {
            auto &matchedMp = mapDB.mapPoints.at(matchedId);
            if (mp.observations.size() < matchedMp.observations.size()) {                  // :510-518
                if (matchedMp.status == Status::NOT_TRIANGULATED) {
                    matchedMp.eraseObservation(kf.id);
                    kf.mapPoints[bestKp.v] = mp.id;
                    mp.addObservation(kf.id, bestKp);
                } else {
                    mp.replaceWith(mapDB, matchedMp);
                }
                erased.insert(mpId);
            } else {                                                                       // :519-522
                matchedMp.replaceWith(mapDB, mp);
                erased.insert(matchedId);
            }
        }
        ++fusedCount;
    }
    return fusedCount;
}

// ---- matchMapPointsSim3 (keyframe_matcher.hpp:85-91, keyframe_matcher.cpp:552-686) --------------------------------
// Stateless: both findMatchesTranformedMps directions and the agreement filter are one sg_match_sim3 call.
template <class Types, class Keyframe, class MpIdT, class MapDB, class Settings>
void projectThroughSim3(const std::vector<MpIdT> &mpIdsA, const std::vector<bool> &alreadyMatchedInA, Keyframe &kfB,
                        const typename Types::Matrix3d &rotBAW, const typename Types::Vector3d &transBAW, MapDB &mapDB,
                        float margin, const Settings &settings, QueryArrays &q) {
    using Status = typename std::decay<decltype(mapDB.mapPoints.begin()->second.status)>::type;
    static const std::uint32_t none[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (std::size_t indA = 0; indA < mpIdsA.size(); ++indA) {
        bool live = false;
        if (!alreadyMatchedInA.at(indA) && mpIdsA[indA].v != -1) {                         // :566-569
            const auto &mp = mapDB.mapPoints.at(mpIdsA[indA]);
            if (mp.status == Status::TRIANGULATED) {                                       // :572
                const typename Types::Vector3d posB = rotBAW * mp.position + transBAW;
                typename Types::Vector2d pix;
                float xRight = 0.0;
                if (reprojectToImage(*kfB.shared->camera, rotBAW, transBAW, mp.position, pix, xRight)) {   // :579-583
                    const double viewingDistance = posB.norm();
                    if (!(viewingDistance < mp.minViewingDistance || mp.maxViewingDistance < viewingDistance)) {   // :590-592
                        const int level = mp.predictScaleLevel(viewingDistance, settings);
                        const auto pixF = pix.template cast<float>();
                        q.push(pixF(0), pixF(1), margin * settings.scaleFactors.at((std::size_t)level), level, mp.descriptor.data());
                        live = true;
                    }
                }
            }
        }
        if (!live) q.push(0.f, 0.f, -1.f, 0, none);                                        // r < 0: no query
    }
}

template <class Types, class Keyframe, class MpIdT, class MapDB, class Settings>
void matchMapPointsSim3(Keyframe &kf1, Keyframe &kf2, const typename Types::Matrix4d &transform12, MapDB &mapDB,
                        std::vector<std::pair<MpIdT, MpIdT>> &matches, const Settings &settings, sg_ctx *ctx) {
    constexpr float margin = 7.5;
    std::vector<bool> already1(kf1.mapPoints.size(), false), already2(kf2.mapPoints.size(), false);
    for (const auto &match : matches) {                                                    // :643-649
        already1.at(mapDB.mapPoints.at(match.first).observations.at(kf1.id).v) = true;
        already2.at(mapDB.mapPoints.at(match.second).observations.at(kf2.id).v) = true;
    }
    auto rot = [](const typename Types::Matrix4d &m) {
        typename Types::Matrix3d r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = m(i, j);
        return r;
    };
    auto trans = [](const typename Types::Matrix4d &m) { return typename Types::Vector3d(m(0, 3), m(1, 3), m(2, 3)); };
    const typename Types::Matrix4d transform21w = transform12.inverse() * kf1.poseCW;      // :651
    const typename Types::Matrix4d transform12W = transform12 * kf2.poseCW;                // :662
    QueryArrays q12, q21;
    projectThroughSim3<Types>(kf1.mapPoints, already1, kf2, rot(transform21w), trans(transform21w), mapDB, margin, settings, q12);
    projectThroughSim3<Types>(kf2.mapPoints, already2, kf1, rot(transform12W), trans(transform12W), mapDB, margin, settings, q21);
    KeypointArrays k1, k2;
    gatherKeypoints(kf1, k1);
    gatherKeypoints(kf2, k2);
    if (k1.n == 0 || k2.n == 0) return;
    std::vector<std::int32_t> pairs(2 * (std::size_t)std::min(k1.n, k2.n));
    std::uint32_t n = 0;
    check(ctx, sg_match_sim3(ctx, k1.x.data(), k1.y.data(), k1.octave.data(), k1.desc.data(), k1.n, k1.order.data(), k2.x.data(),
                             k2.y.data(), k2.octave.data(), k2.desc.data(), k2.n, k2.order.data(), q12.x.data(), q12.y.data(),
                             q12.r.data(), q12.desc.data(), q12.level.data(), q21.x.data(), q21.y.data(), q21.r.data(),
                             q21.desc.data(), q21.level.data(), pairs.data(), &n), "sg_match_sim3");
    for (std::uint32_t i = 0; i < n; ++i)                                                  // :680-682
        matches.emplace_back(kf1.mapPoints.at((std::size_t)pairs[2 * i]), kf2.mapPoints.at((std::size_t)pairs[2 * i + 1]));
}

// ---- matchForTriangulationDBoW (keyframe_matcher.hpp:53, keyframe_matcher.cpp:160-293) ----------------------------
template <class Types, class KpIdT, class Keyframe, class Settings>
std::vector<std::pair<KpIdT, KpIdT>> matchForTriangulationDBoW(Keyframe &kf1, Keyframe &kf2, const Settings &settings, sg_ctx *ctx) {
    auto rot = [](const typename Types::Matrix4d &m) {
        typename Types::Matrix3d r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = m(i, j);
        return r;
    };
    auto trans = [](const typename Types::Matrix4d &m) { return typename Types::Vector3d(m(0, 3), m(1, 3), m(2, 3)); };
    const typename Types::Matrix3d E = Types::createE21(rot(kf2.poseCW), trans(kf2.poseCW), rot(kf1.poseCW), trans(kf1.poseCW));   // :171-175
    sg_triangulation_params tp{};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) tp.E[3 * r + c] = E(r, c);
    tp.scale_factors = settings.scaleFactors.data();
    tp.n_levels = (int)settings.scaleFactors.size();
    tp.residual_deg_thr = settings.parameters.slam.epipolarCheckThresholdDegrees;            // :168
    tp.thr = 50;
    tp.check_orientation = 1;
    struct Side { std::vector<std::uint32_t> desc; std::vector<float> angle; std::vector<std::int32_t> octave, node; std::vector<double> bearing; std::vector<std::uint8_t> elig; };
    auto gather = [](const Keyframe &kf, Side &s) {
        const auto &kps = kf.shared->keyPoints;
        const std::size_t n = kps.size();
        s.desc.resize(8 * n); s.angle.resize(n); s.octave.resize(n); s.node.assign(n, -1); s.bearing.resize(3 * n); s.elig.resize(std::max<std::size_t>(n, 1));
        for (std::size_t i = 0; i < n; ++i) {
            std::memcpy(&s.desc[8 * i], kps[i].descriptor.data(), 32);
            s.angle[i] = kps[i].angle; s.octave[i] = kps[i].octave;
            for (int c = 0; c < 3; ++c) s.bearing[3 * i + c] = kps[i].bearing(c);
            s.elig[i] = kf.mapPoints.at(i).v == -1;                                        // :205-208, :218-221
        }
        for (const auto &e : kf.shared->bowFeatureVec) for (unsigned i : e.second) s.node.at(i) = (std::int32_t)e.first;
    };
    Side a, b;
    gather(kf1, a);
    gather(kf2, b);
    const int nA = (int)a.angle.size(), nB = (int)b.angle.size();
    std::vector<std::pair<KpIdT, KpIdT>> out;
    if (nA == 0 || nB == 0) return out;
    std::vector<std::int32_t> m((std::size_t)nA, -1);
    std::uint32_t n = 0;
    check(ctx, sg_match_triangulation(ctx, a.desc.data(), a.angle.data(), a.octave.data(), a.bearing.data(), a.node.data(), a.elig.data(), nA,
                                      b.desc.data(), b.angle.data(), b.bearing.data(), b.node.data(), b.elig.data(), nB, &tp, m.data(), &n),
          "sg_match_triangulation");
    out.reserve(n);
    for (int i = 0; i < nA; ++i)                                                           // :282-290
        if (m[(std::size_t)i] >= 0) out.emplace_back(KpIdT(i), KpIdT(m[(std::size_t)i]));
    return out;
}

// ---- MapPoint::updateDescriptor (map_point.cpp:75-116) for many map points in one launch ----------------------------
template <class MpIdT, class MapDB>
void updateDescriptors(MapDB &mapDB, const std::vector<MpIdT> &ids, sg_ctx *ctx) {
    std::vector<std::uint32_t> desc;
    std::vector<std::int64_t> offsets(1, 0);
    for (const MpIdT id : ids) {
        const auto &mp = mapDB.mapPoints.at(id);
        for (const auto &obs : mp.observations) {                                          // :78-84
            const auto &kf = *mapDB.keyframes.at(obs.first);
            if (!kf.hasFeatureDescriptors()) continue;
            const auto &d = kf.shared->keyPoints.at((std::size_t)obs.second.v).descriptor;
            desc.insert(desc.end(), d.begin(), d.end());
        }
        offsets.push_back((std::int64_t)(desc.size() / 8));
    }
    if (ids.empty()) return;
    std::vector<std::int32_t> best(ids.size());
    if (desc.empty()) desc.resize(8);
    check(ctx, sg_medoid(ctx, desc.data(), offsets.data(), (int)ids.size(), best.data()), "sg_medoid");
    for (std::size_t s = 0; s < ids.size(); ++s) {
        if (offsets[s + 1] == offsets[s]) continue;                                        // :86 no descriptors: unchanged
        auto &mp = mapDB.mapPoints.at(ids[s]);
        std::memcpy(mp.descriptor.data(), &desc[8 * (std::size_t)(offsets[s] + best[s])], 32);   // :115
    }
}

}  // namespace cuda_matchers
}  // namespace slam
