// The matcher adapters of slam-module_b200/host/slam_matchers.hpp instantiated with THE REFERENCE'S OWN CLASSES
// (slam::Keyframe, slam::MapPoint, slam::MapDB, slam::StaticSettings from /root/reference, compiled verbatim into
// oracle/_ref) and run next to the reference's own functions on identical scenes:
//   searchByProjection, replaceDuplication<vector> / <set>, matchMapPointsSim3, matchForTriangulationDBoW,
//   MapPoint::updateDescriptor.
// After each pair of runs the complete map state (keypoint -> map point tables, every map point's observations,
// status and descriptor, the set of surviving map points) and the return values must be identical.
// TEST INFRASTRUCTURE: needs a GPU (the adapters call libslamgpu.so) and the reference tree at build time
// (oracle/Makefile target `ref_adapter`); exit code 0 = all scenes identical.
#include "../../oracle/ref_slam.cpp"   // scenario builders + the definitions the reference objects link against

#include "../../slam-module_b200/host/slam_matchers.hpp"

namespace {

struct EigenTypes {
    using Vector2f = Eigen::Vector2f;
    using Vector2d = Eigen::Vector2d;
    using Vector3d = Eigen::Vector3d;
    using Matrix3d = Eigen::Matrix3d;
    using Matrix4d = Eigen::Matrix4d;
    static Matrix3d createE21(const Matrix3d &r1, const Vector3d &t1, const Matrix3d &r2, const Vector3d &t2) {
        return openvslam::solve::essential_solver::create_E_21(r1, t1, r2, t2);
    }
};

struct Rng {
    std::uint64_t s;
    explicit Rng(std::uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 1) {}
    std::uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
    double uni() { return (double)(next() >> 11) / 9007199254740992.0; }
    double uni(double a, double b) { return a + (b - a) * uni(); }
    int below(int n) { return (int)(next() % (std::uint64_t)n); }
    double gauss() { double a = 0; for (int i = 0; i < 12; ++i) a += uni(); return a - 6.0; }
    void desc(std::uint32_t *d) { for (int i = 0; i < 8; ++i) d[i] = (std::uint32_t)next(); }
    void flip(std::uint32_t *d, int bits) { for (int i = 0; i < bits; ++i) { const int b = below(256); d[b >> 5] ^= 1u << (b & 31); } }
};

// One synthetic map: 3-D points in front of a 640 x 480 pinhole camera (f = 500), two keyframes looking at them
struct Scene {
    MapDB db;
    std::shared_ptr<Keyframe> kf1, kf2;
    std::vector<MpId> queries;            // map points not yet observed by kf1 (projection / duplication tests)
    std::unique_ptr<SettingsBox> sb;
};

std::shared_ptr<tracker::Camera> camera() {
    auto c = std::make_shared<tracker::Camera>();
    c->fx = c->fy = 500; c->cx = 320; c->cy = 240;
    c->vx0 = 0; c->vy0 = 0; c->vx1 = 640; c->vy1 = 480;
    return c;
}

Eigen::Matrix4d pose(double rx, double ry, double rz, double tx, double ty, double tz) {
    const double cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz), sz = std::sin(rz);
    Eigen::Matrix3d Rx, Ry, Rz;
    Rx << 1, 0, 0, 0, cx, -sx, 0, sx, cx;
    Ry << cy, 0, sy, 0, 1, 0, -sy, 0, cy;
    Rz << cz, -sz, 0, sz, cz, 0, 0, 0, 1;
    const Eigen::Matrix3d R = Rz * Ry * Rx;
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T(i, j) = R(i, j);
    T(0, 3) = tx; T(1, 3) = ty; T(2, 3) = tz;
    return T;
}

// keypoints of a keyframe = projections of the scene points (noisy), descriptors = noisy copies of the point's
void observe(Keyframe &kf, const std::vector<Eigen::Vector3d> &P, const std::vector<std::array<std::uint32_t, 8>> &D, Rng &rng,
             std::vector<int> &pointOfKp) {
    const Eigen::Matrix3d R = kf.poseCW.topLeftCorner<3, 3>();
    const Eigen::Vector3d t = kf.poseCW.block<3, 1>(0, 3);
    kf.shared->keyPoints.clear();
    pointOfKp.clear();
    for (std::size_t k = 0; k < P.size(); ++k) {
        const Eigen::Vector3d pc = R * P[k] + t;
        Eigen::Vector2d pix;
        if (!kf.shared->camera->rayToPixel(pc, pix) || !kf.shared->camera->isValidPixel(pix)) continue;
        if (rng.uni() < 0.1) continue;
        const int copies = rng.uni() < 0.12 ? 2 : 1;     // twins: equal distances inside one radius
        for (int c = 0; c < copies; ++c) {
            KeyPoint kp;
            kp.pt.x = (float)(pix.x() + (c ? 0.5 : rng.gauss() * 0.8));
            kp.pt.y = (float)(std::round((pix.y() + rng.gauss() * 0.8) * 2) / 2);     // half-pixel grid: ties in the Y sort
            kp.angle = (float)rng.uni(0, 360);
            kp.octave = rng.below(8);
            std::memcpy(kp.descriptor.data(), D[k].data(), 32);
            if (!c) rng.flip(kp.descriptor.data(), rng.below(24));
            else std::memcpy(kp.descriptor.data(), kf.shared->keyPoints.back().descriptor.data(), 32), kp.octave = kf.shared->keyPoints.back().octave;
            Eigen::Vector3d ray;
            kf.shared->camera->pixelToRay(Eigen::Vector2d(kp.pt.x, kp.pt.y), ray);
            kp.bearing = ray;
            kf.shared->keyPoints.push_back(kp);
            pointOfKp.push_back((int)k);
        }
    }
    kf.mapPoints.assign(kf.shared->keyPoints.size(), MpId(-1));
    kf.shared->featureSearch = FeatureSearch::create(kf.shared->keyPoints);
}

std::unique_ptr<Scene> makeScene(std::uint64_t seed, int nPoints) {
    Rng rng(seed);
    auto s = std::unique_ptr<Scene>(new Scene());
    s->sb.reset(new SettingsBox(8, 1.2f));
    s->sb->params.slam.epipolarCheckThresholdDegrees = 0.5f;
    std::vector<Eigen::Vector3d> P(nPoints);
    std::vector<std::array<std::uint32_t, 8>> D(nPoints);
    for (int k = 0; k < nPoints; ++k) {
        P[k] = Eigen::Vector3d(rng.uni(-3.5, 3.5), rng.uni(-2.5, 2.5), rng.uni(4, 9));
        rng.desc(D[k].data());
    }
    auto mk = [&](int id, const Eigen::Matrix4d &T) {
        auto kf = std::make_shared<Keyframe>();
        kf->shared = std::make_shared<KeyframeShared>();
        kf->shared->camera = camera();
        kf->id = KfId(id);
        kf->poseCW = T; kf->origPoseCW = T;
        kf->hasFullFeatures = true;
        return kf;
    };
    s->kf1 = mk(1, pose(0.01, -0.02, 0.015, 0.05, -0.02, 0.1));
    s->kf2 = mk(2, pose(-0.015, 0.04, -0.01, -0.35, 0.03, 0.05));
    std::vector<int> p1, p2;
    observe(*s->kf1, P, D, rng, p1);
    observe(*s->kf2, P, D, rng, p2);
    s->db.keyframes.emplace(s->kf1->id, s->kf1);
    s->db.keyframes.emplace(s->kf2->id, s->kf2);
    // vocabulary stand-in: node = a few descriptor bits of the scene point
    for (std::size_t i = 0; i < p1.size(); ++i) s->kf1->shared->bowFeatureVec.addFeature(D[p1[i]][0] % 9u, (unsigned)i);
    for (std::size_t i = 0; i < p2.size(); ++i) s->kf2->shared->bowFeatureVec.addFeature(rng.uni() < 0.9 ? D[p2[i]][0] % 9u : (unsigned)rng.below(9), (unsigned)i);
    // map points: kf2 owns one per keypoint (70 %), kf1 owns some (30 %); the rest of kf2's are the query set for kf1
    int nextId = 0;
    auto newPoint = [&](int k, Keyframe &kf, int kp) -> MapPoint & {
        MapPoint mp(MpId(nextId++), kf.id, KpId(kp));
        mp.position = P[k] + Eigen::Vector3d(rng.gauss() * 0.004, rng.gauss() * 0.004, rng.gauss() * 0.004);
        const Eigen::Vector3d view = (kf.cameraCenter() - mp.position).normalized();
        mp.norm = view.cast<float>();
        if (rng.uni() < 0.3) mp.norm = (view + Eigen::Vector3d(0.25, 0.1, 0)).normalized().cast<float>();    // > 4 degrees off: full radius
        if (rng.uni() < 0.04) mp.norm = Eigen::Vector3f::Zero();
        const double d = (kf.cameraCenter() - mp.position).norm();
        const int oct = kf.shared->keyPoints[kp].octave;
        mp.maxViewingDistance = (float)(d * std::pow(1.2, oct + 0.5));
        mp.minViewingDistance = (float)(d * std::pow(1.2, oct + 0.5) / 3.58);
        if (rng.uni() < 0.04) mp.minViewingDistance = (float)(d * 3);
        const double st = rng.uni();
        mp.status = st < 0.85 ? MapPointStatus::TRIANGULATED : st < 0.93 ? MapPointStatus::NOT_TRIANGULATED : st < 0.97 ? MapPointStatus::UNSURE : MapPointStatus::BAD;
        std::memcpy(mp.descriptor.data(), kf.shared->keyPoints[kp].descriptor.data(), 32);
        rng.flip(mp.descriptor.data(), rng.below(40));
        kf.mapPoints[kp] = mp.id;
        return s->db.mapPoints.emplace(mp.id, mp).first->second;
    };
    for (std::size_t i = 0; i < p2.size(); ++i)
        if (rng.uni() < 0.7) { MapPoint &mp = newPoint(p2[i], *s->kf2, (int)i); s->queries.push_back(mp.id); }
    for (std::size_t i = 0; i < p1.size(); ++i)
        if (rng.uni() < 0.3) {
            MapPoint &mp = newPoint(p1[i], *s->kf1, (int)i);
            // some of kf1's points are seen from further keyframes too (observation counts decide who is replaced)
            const int extra = rng.below(3);
            for (int e = 0; e < extra; ++e) {
                const KfId other(10 + e);
                if (!s->db.keyframes.count(other)) {
                    const std::uint32_t z[8] = {0};
                    s->db.keyframes.emplace(other, make_keyframe(other.v, nullptr, nullptr, nullptr, nullptr, z, nullptr, 1));
                }
                mp.addObservation(other, KpId(0));
            }
        }
    // a few query ids that must be skipped: invalid, duplicated
    s->queries.insert(s->queries.begin() + (long)(s->queries.size() / 2), MpId(-1));
    if (s->queries.size() > 20) s->queries.push_back(s->queries[7]);
    return s;
}

// deep copy of a scene (the reference's Keyframe copy constructor shares `shared`: rebuild explicitly)
std::unique_ptr<Scene> clone(const Scene &a) {
    auto s = std::unique_ptr<Scene>(new Scene());
    s->sb.reset(new SettingsBox(8, 1.2f));
    s->sb->params.slam.epipolarCheckThresholdDegrees = a.sb->params.slam.epipolarCheckThresholdDegrees;
    for (const auto &e : a.db.keyframes) {
        auto kf = std::make_shared<Keyframe>();
        kf->shared = std::make_shared<KeyframeShared>();
        kf->shared->camera = e.second->shared->camera;
        kf->shared->keyPoints = e.second->shared->keyPoints;
        kf->shared->bowFeatureVec = e.second->shared->bowFeatureVec;
        kf->shared->featureSearch = FeatureSearch::create(kf->shared->keyPoints);
        kf->id = e.second->id;
        kf->poseCW = e.second->poseCW; kf->origPoseCW = e.second->origPoseCW;
        kf->mapPoints = e.second->mapPoints;
        kf->hasFullFeatures = true;
        s->db.keyframes.emplace(kf->id, kf);
    }
    s->db.mapPoints = a.db.mapPoints;
    s->kf1 = s->db.keyframes.at(KfId(1));
    s->kf2 = s->db.keyframes.at(KfId(2));
    s->queries = a.queries;
    return s;
}

std::vector<long long> snapshot(const Scene &s) {
    std::vector<long long> v;
    for (const auto &e : s.db.keyframes) {
        v.push_back(1000000 + e.first.v);
        for (const MpId id : e.second->mapPoints) v.push_back(id.v);
    }
    for (const auto &e : s.db.mapPoints) {
        const MapPoint &mp = e.second;
        v.push_back(2000000 + mp.id.v);
        v.push_back((long long)mp.status);
        for (const auto &o : mp.observations) { v.push_back(o.first.v); v.push_back(o.second.v); }
        for (const auto w : mp.descriptor) v.push_back(w);
    }
    return v;
}

int failures = 0;
void expect(bool ok, const char *what, std::uint64_t seed) {
    if (!ok) { std::fprintf(stderr, "MISMATCH: %s (seed %llu)\n", what, (unsigned long long)seed); ++failures; }
}

}  // namespace

int main(int argc, char **argv) {
    const int nPoints = argc > 1 ? std::atoi(argv[1]) : 900;
    sg_params p{};
    p.width = 640; p.height = 480; p.levels = 8; p.scale_factor = 1.2f; p.max_keypoints = 1000;
    p.ini_fast_thr = 20; p.min_fast_thr = 7; p.max_frames = 1;
    sg_ctx *ctx = nullptr;
    if (sg_create(0, &p, &ctx) != SG_OK) { std::fprintf(stderr, "sg_create: %s\n", sg_last_error(nullptr)); return 2; }
    using namespace slam::cuda_matchers;
    long total[6] = {0, 0, 0, 0, 0, 0};
    for (std::uint64_t seed = 1; seed <= 6; ++seed) {
        auto base = makeScene(seed, nPoints);
        {   // searchByProjection: kf2's map points projected into kf1
            auto a = clone(*base), b = clone(*base);
            std::vector<MpId> mps;                       // every valid query id once, in list order
            std::set<int> seen;
            for (const MpId id : a->queries) if (id.v != -1 && seen.insert(id.v).second) mps.push_back(id);
            const int ra = slam::searchByProjection(*a->kf1, mps, a->db, nullptr, 15.0f, a->sb->settings);
            const int rb = searchByProjection<EigenTypes>(*b->kf1, mps, b->db, 15.0f, b->sb->settings, ctx);
            expect(ra == rb, "searchByProjection count", seed);
            expect(snapshot(*a) == snapshot(*b), "searchByProjection state", seed);
            total[0] += ra;
        }
        {   // replaceDuplication over a vector (with an invalid and a duplicated id) and over a set
            auto a = clone(*base), b = clone(*base);
            const unsigned ra = slam::replaceDuplication(*a->kf1, a->queries, 3.0f, a->db, a->sb->settings);
            const unsigned rb = replaceDuplication<EigenTypes>(*b->kf1, b->queries, 3.0f, b->db, b->sb->settings, ctx);
            expect(ra == rb, "replaceDuplication<vector> count", seed);
            expect(snapshot(*a) == snapshot(*b), "replaceDuplication<vector> state", seed);
            total[1] += ra;
            auto c = clone(*base), d = clone(*base);
            std::set<MpId> ids(c->queries.begin(), c->queries.end());
            const unsigned rc = slam::replaceDuplication(*c->kf1, ids, 5.0f, c->db, c->sb->settings);
            const unsigned rd = replaceDuplication<EigenTypes>(*d->kf1, ids, 5.0f, d->db, d->sb->settings, ctx);
            expect(rc == rd, "replaceDuplication<set> count", seed);
            expect(snapshot(*c) == snapshot(*d), "replaceDuplication<set> state", seed);
            total[2] += rc;
        }
        {   // matchMapPointsSim3 with a near-identity Sim3 guess (scale 1.02, small rotation / translation) and seeds
            auto a = clone(*base), b = clone(*base);
            Eigen::Matrix4d T12 = pose(0.004, -0.003, 0.002, 0.01, -0.005, 0.02);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T12(i, j) *= 1.02;
            // transform12 maps kf2 camera coordinates to kf1's: compose with the true relative pose
            const Eigen::Matrix4d rel = a->kf1->poseCW * a->kf2->poseCW.inverse();
            T12 = T12 * rel;
            std::vector<std::pair<MpId, MpId>> ma, mb;
            // seed: the first two (kf1 point, kf2 point) pairs found by a dry run
            {
                auto dry = clone(*base);
                std::vector<std::pair<MpId, MpId>> m0;
                slam::matchMapPointsSim3(*dry->kf1, *dry->kf2, T12, dry->db, m0, dry->sb->settings);
                for (std::size_t i = 0; i < m0.size() && i < 2; ++i) { ma.push_back(m0[i]); mb.push_back(m0[i]); }
            }
            slam::matchMapPointsSim3(*a->kf1, *a->kf2, T12, a->db, ma, a->sb->settings);
            matchMapPointsSim3<EigenTypes>(*b->kf1, *b->kf2, T12, b->db, mb, b->sb->settings, ctx);
            bool same = ma.size() == mb.size();
            for (std::size_t i = 0; same && i < ma.size(); ++i) same = ma[i].first.v == mb[i].first.v && ma[i].second.v == mb[i].second.v;
            expect(same, "matchMapPointsSim3 pairs", seed);
            total[3] += (long)ma.size();
        }
        {   // matchForTriangulationDBoW
            auto a = clone(*base), b = clone(*base);
            const auto ra = slam::matchForTriangulationDBoW(*a->kf1, *a->kf2, a->sb->settings);
            const auto rb = matchForTriangulationDBoW<EigenTypes, KpId>(*b->kf1, *b->kf2, b->sb->settings, ctx);
            bool same = ra.size() == rb.size();
            for (std::size_t i = 0; same && i < ra.size(); ++i) same = ra[i].first.v == rb[i].first.v && ra[i].second.v == rb[i].second.v;
            expect(same, "matchForTriangulationDBoW pairs", seed);
            total[4] += (long)ra.size();
        }
        {   // MapPoint::updateDescriptor: give every map point observations in both keyframes first
            auto a = clone(*base), b = clone(*base);
            std::vector<MpId> ids;
            for (auto *s : {a.get(), b.get()}) {
                Rng r2(seed + 77);
                for (auto &e : s->db.mapPoints) {
                    for (Keyframe *kf : {s->kf1.get(), s->kf2.get()})
                        if (!e.second.observations.count(kf->id) && !kf->shared->keyPoints.empty())
                            e.second.observations.emplace(kf->id, KpId(r2.below((int)kf->shared->keyPoints.size())));
                }
            }
            for (const auto &e : a->db.mapPoints) ids.push_back(e.first);
            for (const MpId id : ids) a->db.mapPoints.at(id).updateDescriptor(a->db);
            updateDescriptors(b->db, ids, ctx);
            expect(snapshot(*a) == snapshot(*b), "updateDescriptor state", seed);
            total[5] += (long)ids.size();
        }
    }
    sg_destroy(ctx);
    std::printf("ref_adapter: searchByProjection %ld matches, replaceDuplication %ld + %ld fused, sim3 %ld pairs, triangulation %ld pairs, "
                "%ld descriptors updated; %d mismatches\n", total[0], total[1], total[2], total[3], total[4], total[5], failures);
    // the scenes must exercise every path
    if (total[0] < 100 || total[1] < 100 || total[2] < 100 || total[3] < 50 || total[4] < 50) { std::fprintf(stderr, "scenes too sparse\n"); return 3; }
    return failures ? 1 : 0;
}
