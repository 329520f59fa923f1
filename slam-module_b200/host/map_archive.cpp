// Map archives: reader / writer of the cereal binary files the reference's Mapper writes (mapper.cpp:504-512) and
// loadMapDB reads (mapper_helpers.cpp:958-993), for the data-model slice of slam_frontend.hpp, and the descriptor
// database rebuilt from a loaded map.
//
// cereal BinaryOutputArchive rules restated (cereal 1.3, portable only between machines of the same endianness):
//   arithmetic / enum           raw bytes of the value (enums: underlying type)
//   std::string / std::vector   element count as uint64, then the elements (arithmetic vectors: one raw block)
//   std::array<arithmetic, N>   one raw block, no count
//   std::map                    element count as uint64, then key, value per entry in key order
//   std::shared_ptr<T>          uint32 id; 0 = null; (id & 0x80000000) = first occurrence, the object follows;
//                               otherwise a reference to the object saved under (id | 0x80000000)
//   classes                     the fields their serialize / save functions list, in order (no version tag is registered
//                               for any of the reference's classes)
// The Eigen and cv::Vec3b adapters live in the parent project (../util/serialization.hpp, absent): see
// MapArchiveOptions in slam_frontend.hpp.  PARITY UNPINNED: no archive written by the reference is available offline.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../include/slamgpu.h"
#include "slam_frontend.hpp"

namespace slam {
namespace {

struct Reader {
    const std::vector<char> &buf;
    size_t at = 0;
    bool ok = true;
    std::string err;
    MapArchiveOptions opt;
    std::map<std::uint32_t, std::shared_ptr<void>> shared;
    explicit Reader(const std::vector<char> &b) : buf(b) {}
    void fail(const char *what) { if (ok) { ok = false; err = std::string(what) + " at byte " + std::to_string(at); } }
    void raw(void *dst, size_t n) {
        if (!ok) return;
        if (at + n > buf.size()) { fail("archive truncated"); return; }
        std::memcpy(dst, buf.data() + at, n);
        at += n;
    }
    template <class T> T pod() { T v{}; raw(&v, sizeof(T)); return v; }
    size_t count(size_t elemBytes) {
        const std::uint64_t n = pod<std::uint64_t>();
        if (ok && n * std::max<size_t>(elemBytes, 1) > buf.size() - at) { fail("element count beyond the end of the archive"); return 0; }
        return (size_t)n;
    }
    template <class M> void matrix(M &m, int rows, int cols) {   // fixed size, column major in the file
        if (opt.fixedSizeHeader) {
            const std::int32_t r = pod<std::int32_t>(), c = pod<std::int32_t>();
            if (ok && (r != rows || c != cols)) fail("unexpected matrix extent");
        }
        for (int c = 0; c < cols; ++c)
            for (int r = 0; r < rows; ++r) m(r, c) = pod<typename std::decay<decltype(m(0, 0))>::type>();
    }
    template <class V> void vector3(V &v) {
        if (opt.fixedSizeHeader) { pod<std::int32_t>(); pod<std::int32_t>(); }
        for (int i = 0; i < 3; ++i) v(i) = pod<typename std::decay<decltype(v(0))>::type>();
    }
};

struct Writer {
    std::vector<char> buf;
    MapArchiveOptions opt;
    std::map<const void *, std::uint32_t> shared;
    std::uint32_t nextId = 1;
    void raw(const void *src, size_t n) { const char *p = (const char *)src; buf.insert(buf.end(), p, p + n); }
    template <class T> void pod(const T &v) { raw(&v, sizeof(T)); }
    void count(size_t n) { pod<std::uint64_t>((std::uint64_t)n); }
    template <class M> void matrix(const M &m, int rows, int cols) {
        if (opt.fixedSizeHeader) { pod<std::int32_t>(rows); pod<std::int32_t>(cols); }
        for (int c = 0; c < cols; ++c)
            for (int r = 0; r < rows; ++r) pod(m(r, c));
    }
    template <class V> void vector3(const V &v) {
        if (opt.fixedSizeHeader) { pod<std::int32_t>(3); pod<std::int32_t>(1); }
        for (int i = 0; i < 3; ++i) pod(v(i));
    }
};

// ---- KeyPoint (key_point.hpp:22-25): pt.x, pt.y, angle, octave, octave, bearing, descriptor -------------------------
void load(Reader &r, KeyPoint &kp) {
    kp.pt.x = r.pod<float>(); kp.pt.y = r.pod<float>(); kp.angle = r.pod<float>();
    kp.octave = r.pod<std::int32_t>();
    kp.octave = r.pod<std::int32_t>();      // written twice; the second value wins on load, as in cereal
    r.vector3(kp.bearing);
    r.raw(kp.descriptor.data(), 32);
}
void save(Writer &w, const KeyPoint &kp) {
    w.pod(kp.pt.x); w.pod(kp.pt.y); w.pod(kp.angle);
    w.pod<std::int32_t>(kp.octave); w.pod<std::int32_t>(kp.octave);
    w.vector3(kp.bearing);
    w.raw(kp.descriptor.data(), 32);
}

// ---- KeyframeShared (keyframe.hpp:80-105): cameraModel, keyPoints, colors, stereoPointCloud -------------------------
void load(Reader &r, KeyframeShared &s) {
    const size_t nc = r.count(1);
    s.cameraModel.resize(nc);
    r.raw(&s.cameraModel[0], nc);
    const size_t nk = r.count(4 * 5 + 24 + 32);
    s.keyPoints.resize(nk);
    for (auto &kp : s.keyPoints) load(r, kp);
    const size_t ncol = r.count(3);
    s.colors.resize(ncol);
    for (auto &c : s.colors) r.raw(c.data(), 3);
    const std::uint32_t id = r.pod<std::uint32_t>();
    s.stereoPointCloud.reset();
    if (id & 0x80000000u) {
        auto cloud = std::make_shared<std::vector<Vector3f>>(r.count(12));
        for (auto &p : *cloud) r.vector3(p);
        r.shared[id & 0x7fffffffu] = cloud;
        s.stereoPointCloud = cloud;
    } else if (id) {
        auto it = r.shared.find(id);
        if (it == r.shared.end()) r.fail("dangling shared pointer id");
        else s.stereoPointCloud = std::static_pointer_cast<std::vector<Vector3f>>(it->second);
    }
    auto cam = std::make_shared<tracker::Camera>();       // tracker::Camera::deserialize(cameraModel) belongs to the parent project
    s.camera = cam;
}
void save(Writer &w, const KeyframeShared &s) {
    w.count(s.cameraModel.size());
    w.raw(s.cameraModel.data(), s.cameraModel.size());
    w.count(s.keyPoints.size());
    for (const auto &kp : s.keyPoints) save(w, kp);
    w.count(s.colors.size());
    for (const auto &c : s.colors) w.raw(c.data(), 3);
    if (!s.stereoPointCloud) { w.pod<std::uint32_t>(0); return; }
    auto it = w.shared.find(s.stereoPointCloud.get());
    if (it != w.shared.end()) { w.pod<std::uint32_t>(it->second); return; }
    const std::uint32_t id = w.nextId++;
    w.shared[s.stereoPointCloud.get()] = id;
    w.pod<std::uint32_t>(id | 0x80000000u);
    w.count(s.stereoPointCloud->size());
    for (const auto &p : *s.stereoPointCloud) w.vector3(p);
}

// ---- Keyframe (keyframe.hpp:199-213) ---------------------------------------------------------------------------------
void load(Reader &r, Keyframe &kf) {
    const std::uint32_t id = r.pod<std::uint32_t>();
    if (id & 0x80000000u) {
        auto sh = std::make_shared<KeyframeShared>();
        load(r, *sh);
        r.shared[id & 0x7fffffffu] = sh;
        kf.shared = sh;
    } else if (id) {
        auto it = r.shared.find(id);
        if (it == r.shared.end()) r.fail("dangling shared pointer id");
        else kf.shared = std::static_pointer_cast<KeyframeShared>(it->second);
    } else {
        kf.shared.reset();
    }
    kf.id.v = r.pod<std::int32_t>(); kf.previousKfId.v = r.pod<std::int32_t>(); kf.nextKfId.v = r.pod<std::int32_t>();
    const size_t nt = r.count(8);
    kf.keyPointToTrackId.clear();
    for (size_t i = 0; i < nt && r.ok; ++i) { const int k = r.pod<std::int32_t>(), t = r.pod<std::int32_t>(); kf.keyPointToTrackId.emplace(KpId(k), TrackId(t)); }
    kf.mapPoints.resize(r.count(4));
    for (auto &m : kf.mapPoints) m.v = r.pod<std::int32_t>();
    kf.keyPointDepth.resize(r.count(4));
    r.raw(kf.keyPointDepth.data(), 4 * kf.keyPointDepth.size());
    r.matrix(kf.poseCW, 4, 4); r.matrix(kf.origPoseCW, 4, 4); r.matrix(kf.uncertainty, 3, 6);
    kf.t = r.pod<double>();
    kf.hasFullFeatures = r.pod<std::uint8_t>() != 0;
}
void save(Writer &w, const Keyframe &kf) {
    if (!kf.shared) w.pod<std::uint32_t>(0);
    else {
        auto it = w.shared.find(kf.shared.get());
        if (it != w.shared.end()) w.pod<std::uint32_t>(it->second);
        else {
            const std::uint32_t id = w.nextId++;
            w.shared[kf.shared.get()] = id;
            w.pod<std::uint32_t>(id | 0x80000000u);
            save(w, *kf.shared);
        }
    }
    w.pod<std::int32_t>(kf.id.v); w.pod<std::int32_t>(kf.previousKfId.v); w.pod<std::int32_t>(kf.nextKfId.v);
    w.count(kf.keyPointToTrackId.size());
    for (const auto &e : kf.keyPointToTrackId) { w.pod<std::int32_t>(e.first.v); w.pod<std::int32_t>(e.second.v); }
    w.count(kf.mapPoints.size());
    for (const auto &m : kf.mapPoints) w.pod<std::int32_t>(m.v);
    w.count(kf.keyPointDepth.size());
    w.raw(kf.keyPointDepth.data(), 4 * kf.keyPointDepth.size());
    w.matrix(kf.poseCW, 4, 4); w.matrix(kf.origPoseCW, 4, 4); w.matrix(kf.uncertainty, 3, 6);
    w.pod(kf.t);
    w.pod<std::uint8_t>(kf.hasFullFeatures ? 1 : 0);
}

// ---- MapPoint (map_point.hpp:78-93) ----------------------------------------------------------------------------------
void load(Reader &r, MapPoint &mp) {
    mp.id.v = r.pod<std::int32_t>(); mp.trackId.v = r.pod<std::int32_t>();
    mp.status = (MapPointStatus)r.pod<std::int32_t>();
    r.vector3(mp.position); r.vector3(mp.norm);
    mp.minViewingDistance = r.pod<float>(); mp.maxViewingDistance = r.pod<float>();
    r.raw(mp.descriptor.data(), 32);
    const size_t no = r.count(8);
    mp.observations.clear();
    for (size_t i = 0; i < no && r.ok; ++i) { const int k = r.pod<std::int32_t>(), p = r.pod<std::int32_t>(); mp.observations.emplace(KfId(k), KpId(p)); }
    mp.referenceKeyframe.v = r.pod<std::int32_t>();
    r.raw(mp.color.data(), 3);
}
void save(Writer &w, const MapPoint &mp) {
    w.pod<std::int32_t>(mp.id.v); w.pod<std::int32_t>(mp.trackId.v); w.pod<std::int32_t>((std::int32_t)mp.status);
    w.vector3(mp.position); w.vector3(mp.norm);
    w.pod(mp.minViewingDistance); w.pod(mp.maxViewingDistance);
    w.raw(mp.descriptor.data(), 32);
    w.count(mp.observations.size());
    for (const auto &e : mp.observations) { w.pod<std::int32_t>(e.first.v); w.pod<std::int32_t>(e.second.v); }
    w.pod<std::int32_t>(mp.referenceKeyframe.v);
    w.raw(mp.color.data(), 3);
}

bool fail(std::string *error, const std::string &what) { if (error) *error = what; return false; }

}  // namespace

// ---- MapDB (mapdb.hpp:83-98) -----------------------------------------------------------------------------------------
bool loadMapArchive(const std::string &path, MapDB &db, std::string *error, const MapArchiveOptions &opt) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return fail(error, "cannot open " + path);
    std::vector<char> buf((size_t)f.tellg());
    f.seekg(0);
    f.read(buf.data(), (std::streamsize)buf.size());
    Reader r(buf);
    r.opt = opt;
    db = MapDB();
    const size_t nkf = r.count(4);
    for (size_t i = 0; i < nkf && r.ok; ++i) {
        const KfId key(r.pod<std::int32_t>());
        const std::uint32_t id = r.pod<std::uint32_t>();
        std::shared_ptr<Keyframe> kf;
        if (id & 0x80000000u) {
            kf = std::make_shared<Keyframe>();
            load(r, *kf);
            r.shared[id & 0x7fffffffu] = kf;
        } else if (id) {
            auto it = r.shared.find(id);
            if (it == r.shared.end()) r.fail("dangling shared pointer id");
            else kf = std::static_pointer_cast<Keyframe>(it->second);
        }
        db.keyframes.emplace(key, kf);
    }
    const size_t nmp = r.count(4);
    for (size_t i = 0; i < nmp && r.ok; ++i) {
        const MpId key(r.pod<std::int32_t>());
        MapPoint mp;
        load(r, mp);
        db.mapPoints.emplace(key, mp);
    }
    const size_t ntr = r.count(8);
    for (size_t i = 0; i < ntr && r.ok; ++i) { const int t = r.pod<std::int32_t>(), m = r.pod<std::int32_t>(); db.trackIdToMapPoint.emplace(TrackId(t), MpId(m)); }
    db.loopClosureEdges.resize(r.count(8));
    for (auto &e : db.loopClosureEdges) { e.kfId1.v = r.pod<std::int32_t>(); e.kfId2.v = r.pod<std::int32_t>(); r.matrix(e.poseDiff, 4, 4); }
    r.matrix(db.prevPose, 4, 4); r.matrix(db.prevInputPose, 4, 4);
    {
        const std::int64_t rows = r.pod<std::int64_t>(), cols = r.pod<std::int64_t>();
        if (r.ok && (rows < 0 || cols < 0 || (std::uint64_t)rows * (std::uint64_t)cols * 8 > buf.size())) r.fail("bad dynamic matrix extent");
        db.discardedUncertaintyRows = (int)rows; db.discardedUncertaintyCols = (int)cols;
        db.discardedUncertainty.assign(r.ok ? (size_t)(rows * cols) : 0, 0.0);
        r.raw(db.discardedUncertainty.data(), 8 * db.discardedUncertainty.size());
    }
    db.firstKfTimestamp = r.pod<double>();
    db.nextMp = r.pod<std::int32_t>();
    db.lastKfCandidateId.v = r.pod<std::int32_t>();
    db.lastKfId.v = r.pod<std::int32_t>();
    if (r.ok && r.at != buf.size()) r.fail("trailing bytes");
    if (!r.ok) return fail(error, r.err);
    return true;
}

bool saveMapArchive(const std::string &path, const MapDB &db, std::string *error, const MapArchiveOptions &opt) {
    Writer w;
    w.opt = opt;
    w.count(db.keyframes.size());
    for (const auto &e : db.keyframes) {
        w.pod<std::int32_t>(e.first.v);
        if (!e.second) { w.pod<std::uint32_t>(0); continue; }
        auto it = w.shared.find(e.second.get());
        if (it != w.shared.end()) { w.pod<std::uint32_t>(it->second); continue; }
        const std::uint32_t id = w.nextId++;
        w.shared[e.second.get()] = id;
        w.pod<std::uint32_t>(id | 0x80000000u);
        save(w, *e.second);
    }
    w.count(db.mapPoints.size());
    for (const auto &e : db.mapPoints) { w.pod<std::int32_t>(e.first.v); save(w, e.second); }
    w.count(db.trackIdToMapPoint.size());
    for (const auto &e : db.trackIdToMapPoint) { w.pod<std::int32_t>(e.first.v); w.pod<std::int32_t>(e.second.v); }
    w.count(db.loopClosureEdges.size());
    for (const auto &e : db.loopClosureEdges) { w.pod<std::int32_t>(e.kfId1.v); w.pod<std::int32_t>(e.kfId2.v); w.matrix(e.poseDiff, 4, 4); }
    w.matrix(db.prevPose, 4, 4); w.matrix(db.prevInputPose, 4, 4);
    w.pod<std::int64_t>(db.discardedUncertaintyRows); w.pod<std::int64_t>(db.discardedUncertaintyCols);
    w.raw(db.discardedUncertainty.data(), 8 * db.discardedUncertainty.size());
    w.pod(db.firstKfTimestamp);
    w.pod<std::int32_t>(db.nextMp);
    w.pod<std::int32_t>(db.lastKfCandidateId.v);
    w.pod<std::int32_t>(db.lastKfId.v);
    std::ofstream f(path, std::ios::binary);
    if (!f) return fail(error, "cannot open " + path + " for writing");
    f.write(w.buf.data(), (std::streamsize)w.buf.size());
    return (bool)f;
}

// ---- rebuild hook: descriptor database of a loaded map ---------------------------------------------------------------
::sg_db *buildDescriptorDatabase(MapDB &mapDB, sg_ctx *ctx, std::vector<KfId> &keyframeIds) {
    keyframeIds.clear();
    std::vector<std::uint32_t> desc;
    std::vector<float> angle;
    std::vector<std::int64_t> offsets(1, 0);
    for (auto &e : mapDB.keyframes) {
        if (!e.second || !e.second->shared) continue;
        KeyframeShared &s = *e.second->shared;
        s.featureSearch = FeatureSearch::create(s.keyPoints);          // mapper_helpers.cpp:982-988
        for (const auto &kp : s.keyPoints) {
            desc.insert(desc.end(), kp.descriptor.begin(), kp.descriptor.end());
            angle.push_back(kp.angle);
        }
        offsets.push_back((std::int64_t)angle.size());
        keyframeIds.push_back(e.first);
    }
    ::sg_db *db = nullptr;
    if (keyframeIds.empty()) return nullptr;
    if (desc.empty()) { desc.resize(8); angle.resize(1); }
    const int rc = sg_db_create(ctx, desc.data(), angle.data(), offsets.data(), (int)keyframeIds.size(), &db);
    if (rc != SG_OK) { std::fprintf(stderr, "slam-b200: sg_db_create failed (%d): %s\n", rc, sg_last_error(ctx)); std::abort(); }
    return db;
}

}  // namespace slam
