// Drives the C++ adapters (slam-module_b200/host/slam_frontend.hpp) the way the reference's callers drive
// the original classes (keyframe.cpp:95-116, mapper_helpers.cpp:1190-1194, loop_closer.cpp:195) and dumps
// every artefact as raw binary; tests/test_gpu_host_adapters.py compares the dumps with the CPU oracle.
//
//   host_adapter_main <w> <h> <imgA.raw> <imgB.raw> <outdir> <maxKeypoints>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <set>
#include <string>
#include <vector>

#include "../../slam-module_b200/host/slam_frontend.hpp"

using namespace slam;

static std::vector<std::uint8_t> readRaw(const char *path, size_t n) {
    std::vector<std::uint8_t> v(n);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)n);
    if ((size_t)f.gcount() != n) { std::fprintf(stderr, "short read %s\n", path); std::exit(2); }
    return v;
}
template <class T>
static std::vector<T> readAll(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return {};
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v((size_t)bytes / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), bytes);
    return v;
}
template <class T>
static void dump(const std::string &path, const std::vector<T> &v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}
static void dumpKeypoints(const std::string &prefix, const KeyPointVector &kps) {
    std::vector<float> xya;
    std::vector<std::int32_t> oct;
    std::vector<std::uint32_t> desc;
    for (const auto &kp : kps) {
        xya.push_back(kp.pt.x); xya.push_back(kp.pt.y); xya.push_back(kp.angle);
        oct.push_back(kp.octave);
        desc.insert(desc.end(), kp.descriptor.begin(), kp.descriptor.end());
    }
    dump(prefix + "_xya.f32", xya);
    dump(prefix + "_octave.i32", oct);
    dump(prefix + "_desc.u32", desc);
}

// a camera whose right-most 40 columns are invalid (exercises dropInvalidKeypoints, orb_extractor.cpp:221-237)
struct CroppedCamera : tracker::Camera {
    double maxX;
    explicit CroppedCamera(double m) : maxX(m) {}
    bool isValidPixel(double x, double y) const override { (void)y; return x < maxX; }
};

int main(int argc, char **argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: %s w h imgA imgB outdir maxKeypoints\n", argv[0]); return 2; }
    const int w = std::atoi(argv[1]), h = std::atoi(argv[2]);
    const std::string out = argv[5];
    auto a = readRaw(argv[3], (size_t)w * h), b = readRaw(argv[4], (size_t)w * h);

    odometry::Parameters params;
    params.slam.maxKeypoints = (unsigned)std::atoi(argv[6]);
    params.slam.cudaMaxFrames = 2;
    params.slam.cudaMaxTracks = 64;
    params.slam.orbLkTrackLevel = 1;
    StaticSettings settings(params);
    tracker::Image imgA(a.data(), w, h, w), imgB(b.data(), w, h, w);

    // ---- pyramid + detector used separately, as feature_detector.cpp / orb_extractor.cpp do -----------
    auto pyramid = ImagePyramid::build(settings, imgA);
    auto detector = FeatureDetector::build(settings, imgA);
    pyramid->update(imgA);
    std::vector<std::int32_t> dims;
    for (size_t l = 0; l < pyramid->numberOfLevels(); ++l) {
        accelerated::Image &lv = pyramid->getLevel(l), &bl = pyramid->getBlurredLevel(l);
        dims.push_back(lv.width); dims.push_back(lv.height);
        dump(out + "/pyr_" + std::to_string(l) + ".u8", std::vector<std::uint8_t>(lv.data, lv.data + (size_t)lv.width * lv.height));
        dump(out + "/blur_" + std::to_string(l) + ".u8", std::vector<std::uint8_t>(bl.data, bl.data + (size_t)bl.width * bl.height));
        if (pyramid->getGpuLevel(l).storageType != accelerated::Image::StorageType::GPU) return 3;
    }
    dump(out + "/dims.i32", dims);
    std::vector<KeyPointVector> perLevel;
    const size_t total = detector->detect(*pyramid, perLevel);
    std::vector<std::int32_t> det;   // level, x, y triples
    size_t seen = 0;
    for (size_t l = 0; l < perLevel.size(); ++l)
        for (const auto &kp : perLevel[l]) {
            det.push_back((int)l); det.push_back((int)kp.pt.x); det.push_back((int)kp.pt.y);
            if (kp.octave != (int)l || kp.angle != 0) return 4;
            ++seen;
        }
    if (seen != total) return 5;
    dump(out + "/detect.i32", det);

    // ---- extractor with tracker features and a camera that rejects part of the image ----------------
    auto extractor = OrbExtractor::build(settings);
    std::vector<tracker::Feature> tracks;
    for (int i = 0; i < 40; ++i) {
        tracker::Feature t;
        t.id = 100 + i;
        t.points[0] = {(float)(7 + (i * 37) % (w - 3)) + 0.25f * (i % 4), (float)(5 + (i * 53) % (h - 3)) + 0.5f * (i % 2)};
        tracks.push_back(t);
    }
    CroppedCamera cam(w - 40.0);
    KeyPointVector kpsA, kpsB;
    std::vector<int> idsA, idsB;
    extractor->detectAndExtract(imgA, cam, tracks, kpsA, idsA);
    dumpKeypoints(out + "/kpsA", kpsA);
    dump(out + "/idsA.i32", std::vector<std::int32_t>(idsA.begin(), idsA.end()));
    tracker::Camera all;
    extractor->detectAndExtract(imgB, all, {}, kpsB, idsB);
    dumpKeypoints(out + "/kpsB", kpsB);

    // batched form must agree with the single-frame calls
    std::vector<KeyPointVector> batch;
    extractor->detectAndExtractBatch({&imgB, &imgA}, all, batch);
    dumpKeypoints(out + "/batch0", batch[0]);
    dumpKeypoints(out + "/batch1", batch[1]);

    // ---- loop-closure matcher on two keyframes; every third map point of kf2 is not triangulated -----
    Keyframe kf1, kf2;
    kf1.shared = std::make_shared<KeyframeShared>(); kf2.shared = std::make_shared<KeyframeShared>();
    kf1.shared->keyPoints = batch[1]; kf2.shared->keyPoints = kpsB;
    MapDB db1, db2;
    for (size_t i = 0; i < kf1.shared->keyPoints.size(); ++i) {
        MpId id; id.v = (i % 5 == 4) ? -1 : (int)db1.mapPoints.size();
        if (id.v >= 0) { MapPoint mp; mp.id = id; mp.status = (i % 7 == 6) ? MapPointStatus::NOT_TRIANGULATED : MapPointStatus::TRIANGULATED; db1.mapPoints.emplace(id, mp); }
        kf1.mapPoints.push_back(id);
    }
    for (size_t i = 0; i < kf2.shared->keyPoints.size(); ++i) {
        MpId id; id.v = (int)db2.mapPoints.size();
        MapPoint mp; mp.id = id; mp.status = (i % 3 == 2) ? MapPointStatus::NOT_TRIANGULATED : MapPointStatus::TRIANGULATED;
        db2.mapPoints.emplace(id, mp);
        kf2.mapPoints.push_back(id);
    }
    auto fe = cudaFrontend(settings, w, h);
    std::vector<int> matched;
    // empty feature vectors: the reference's node walk finds nothing (keyframe_matcher.cpp:70)
    if (matchForLoopClosures(kf1, kf2, db1, db2, matched, params.slam, cudaContext(fe)) != 0) return 6;
    // one node holding every feature: the brute-force degenerate case
    for (size_t i = 0; i < kf1.shared->keyPoints.size(); ++i) kf1.shared->bowFeatureVec[0].push_back((unsigned)i);
    for (size_t i = 0; i < kf2.shared->keyPoints.size(); ++i) kf2.shared->bowFeatureVec[0].push_back((unsigned)i);
    matched.clear();
    const unsigned n = matchForLoopClosures(kf1, kf2, db1, db2, matched, params.slam, cudaContext(fe));
    kf1.shared->bowFeatureVec.clear();
    kf2.shared->bowFeatureVec.clear();
    std::vector<std::int32_t> m(matched.begin(), matched.end());
    m.push_back((int)n);
    dump(out + "/loop_matches.i32", m);

    // the same keyframes with DBoW2-style feature vectors: node = a few descriptor bits (vocabulary stand-in)
    for (size_t i = 0; i < kf1.shared->keyPoints.size(); ++i)
        kf1.shared->bowFeatureVec[(kf1.shared->keyPoints[i].descriptor[0] ^ kf1.shared->keyPoints[i].descriptor[3]) % 7u].push_back((unsigned)i);
    for (size_t i = 0; i < kf2.shared->keyPoints.size(); ++i)
        kf2.shared->bowFeatureVec[(kf2.shared->keyPoints[i].descriptor[0] ^ kf2.shared->keyPoints[i].descriptor[3]) % 7u].push_back((unsigned)i);
    std::vector<int> matchedBow;
    const unsigned nbow = matchForLoopClosures(kf1, kf2, db1, db2, matchedBow, params.slam, cudaContext(fe));
    std::vector<std::int32_t> mw(matchedBow.begin(), matchedBow.end());
    mw.push_back((int)nbow);
    dump(out + "/loop_matches_bow.i32", mw);

    std::vector<int> bf;
    const unsigned nb = bruteForceMatch(batch[1], kpsB, bf, 0.8f, true, cudaContext(fe));
    std::vector<std::int32_t> mb(bf.begin(), bf.end());
    mb.push_back((int)nb);
    dump(out + "/bf_matches.i32", mb);

    std::vector<unsigned> dist(std::min(kpsA.size(), kpsB.size()));
    std::vector<std::uint32_t> da, dbv;
    for (size_t i = 0; i < dist.size(); ++i) {
        da.insert(da.end(), kpsA[i].descriptor.begin(), kpsA[i].descriptor.end());
        dbv.insert(dbv.end(), kpsB[i].descriptor.begin(), kpsB[i].descriptor.end());
    }
    match::compute_descriptor_distance_32(da.data(), dbv.data(), (int)dist.size(), dist.data(), cudaContext(fe));
    dump(out + "/hamming.u32", std::vector<std::uint32_t>(dist.begin(), dist.end()));
    // ---- BowIndex: transform / add / remove / getBowSimilar with the vocabulary the test wrote (bow_index.hpp) ----
    BowVocabulary voc;
    voc.childOff = readAll<std::int32_t>(out + "/voc_child_off.i32");
    if (!voc.childOff.empty()) {
        voc.childIds = readAll<std::int32_t>(out + "/voc_child_ids.i32");
        voc.nodeWord = readAll<std::int32_t>(out + "/voc_node_word.i32");
        voc.nodeDescriptor = readAll<std::uint32_t>(out + "/voc_node_desc.u32");
        voc.nodeWeight = readAll<double>(out + "/voc_node_weight.f64");
        voc.levels = readAll<std::int32_t>(out + "/voc_levels.i32").at(0);
        BowIndex bow(params.slam, voc, cudaContext(fe), 64);
        // keyframes 1..8: A, B, B again (a twin), and slices of both
        std::vector<Keyframe> kfs;
        auto slice = [](const KeyPointVector &v, size_t lo, size_t hi) { return KeyPointVector(v.begin() + (long)lo, v.begin() + (long)hi); };
        const std::vector<KeyPointVector> sets = {batch[1], kpsB, batch[0], slice(batch[1], 0, batch[1].size() / 2),
                                                  slice(kpsB, kpsB.size() / 3, kpsB.size()), slice(batch[1], batch[1].size() / 4, batch[1].size()),
                                                  slice(kpsB, 0, 50), KeyPointVector()};
        for (size_t i = 0; i < sets.size(); ++i) {
            Keyframe kf;
            kf.id.v = (int)i + 1;
            kf.shared = std::make_shared<KeyframeShared>();
            kf.shared->keyPoints = sets[i];
            bow.transform(kf.shared->keyPoints, kf.shared->bowVec, kf.shared->bowFeatureVec);
            MapId mapId; mapId.v = (int)(i % 2);
            bow.add(kf, mapId);
            kfs.push_back(kf);
        }
        std::vector<std::uint32_t> vw, fvNode, fvFeat;
        std::vector<double> vv;
        for (const auto &e : kfs[0].shared->bowVec) { vw.push_back(e.first); vv.push_back(e.second); }
        for (const auto &e : kfs[0].shared->bowFeatureVec)
            for (unsigned f : e.second) { fvNode.push_back(e.first); fvFeat.push_back(f); }
        dump(out + "/bow_vec_word.u32", vw);
        dump(out + "/bow_vec_value.f64", vv);
        dump(out + "/bow_fv_node.u32", fvNode);
        dump(out + "/bow_fv_feature.u32", fvFeat);
        MapKf gone; gone.mapId.v = 1; gone.kfId.v = 6;
        bow.remove(gone);
        std::vector<std::int32_t> simIds;     // per query: count, then (map, kf) pairs
        std::vector<float> simScores;
        for (int q : {0, 1, 4}) {
            const auto similar = bow.getBowSimilar(db1, Atlas(), kfs[(size_t)q]);
            simIds.push_back((int)similar.size());
            for (const auto &sim : similar) { simIds.push_back(sim.mapKf.mapId.v); simIds.push_back(sim.mapKf.kfId.v); simScores.push_back(sim.score); }
        }
        dump(out + "/bow_similar_ids.i32", simIds);
        dump(out + "/bow_similar_scores.f32", simScores);
    }
    // ---- candidate-list matchers on the adapter's own Keyframe / MapPoint / MapDB types ------------------------------
    // (exact parity of these templates is checked against the reference's functions in tests/cpp/ref_adapter_main.cpp;
    //  here: the instantiations exported by libslam_frontend.so run and their results satisfy the matchers' invariants)
    {
        auto cam = std::make_shared<tracker::Camera>();          // f = 1, c = 0: pixel = position.xy / position.z
        auto mkKf = [&](int id, const KeyPointVector &kps) {
            auto kf = std::make_shared<Keyframe>();
            kf->id = KfId(id);
            kf->shared = std::make_shared<KeyframeShared>();
            kf->shared->camera = cam;
            kf->shared->keyPoints = kps;
            kf->shared->featureSearch = FeatureSearch::create(kps);
            kf->mapPoints.assign(kps.size(), MpId(-1));
            return kf;
        };
        MapDB db;
        auto k1 = mkKf(1, batch[1]), k2 = mkKf(2, batch[1]);      // two views of the same keypoints
        db.keyframes.emplace(k1->id, k1);
        db.keyframes.emplace(k2->id, k2);
        std::vector<MpId> ids;
        const auto &kps = k2->shared->keyPoints;
        for (size_t i = 0; i < kps.size(); i += 2) {               // kf2 owns a map point on every second keypoint
            MapPoint mp(MpId((int)ids.size()), k2->id, KpId((int)i));
            mp.status = MapPointStatus::TRIANGULATED;
            mp.position = Vector3d(kps[i].pt.x * 2.0, kps[i].pt.y * 2.0, 2.0);
            mp.norm = (-mp.position).normalized().cast<float>();
            const double d = mp.position.norm();
            mp.maxViewingDistance = (float)(d * std::pow(1.2, kps[i].octave + 0.5));
            mp.minViewingDistance = (float)(d * 0.1);
            mp.descriptor = kps[i].descriptor;
            mp.descriptor[3] ^= 0x00010010u;                        // two flipped bits
            db.mapPoints.emplace(mp.id, mp);
            k2->mapPoints[i] = mp.id;
            ids.push_back(mp.id);
        }
        sg_ctx *c = cudaContext(fe);
        const int found = searchByProjection(*k1, ids, db, nullptr, 15.0f, settings, c);
        int good = 0;
        for (size_t i = 0; i < kps.size(); ++i) {
            const MpId id = k1->mapPoints[i];
            if (id.v < 0) continue;
            const MapPoint &mp = db.mapPoints.at(id);
            if (!mp.observations.count(k1->id) || mp.observations.at(k1->id).v != (int)i) return 7;
            unsigned dist = 0;
            for (int wd = 0; wd < 8; ++wd) dist += (unsigned)__builtin_popcount(mp.descriptor[wd] ^ kps[i].descriptor[wd]);
            if (dist > HAMMING_DIST_THR_HIGH) return 8;
            ++good;
        }
        if (good != found || found < (int)ids.size() / 2) return 9;
        // every point is now seen by kf1: replaceDuplication must skip them all (keyframe_matcher.cpp:429-431) ...
        if (replaceDuplication(*k1, ids, 3.0f, db, settings, c) != 0) return 10;
        // ... and fuse them again after kf1 forgets them
        for (const MpId id : ids) {
            MapPoint &mp = db.mapPoints.at(id);
            if (mp.observations.count(k1->id)) { k1->mapPoints[mp.observations.at(k1->id).v] = MpId(-1); mp.eraseObservation(k1->id); }
        }
        const unsigned fused = replaceDuplication(*k1, std::set<MpId>(ids.begin(), ids.end()), 3.0f, db, settings, c);
        if ((int)fused < found / 2) return 11;
        // Sim3 with the identity guess: every point seen by both keyframes at the same pixel agrees in both directions
        std::vector<std::pair<MpId, MpId>> pairs;
        MapDB db2;
        auto s1 = mkKf(1, batch[1]), s2 = mkKf(2, batch[1]);
        db2.keyframes.emplace(s1->id, s1);
        db2.keyframes.emplace(s2->id, s2);
        int next = 0;
        for (size_t i = 0; i < kps.size(); i += 3)
            for (auto *kf : {s1.get(), s2.get()}) {
                MapPoint mp(MpId(next++), kf->id, KpId((int)i));
                mp.status = MapPointStatus::TRIANGULATED;
                mp.position = Vector3d(kps[i].pt.x * 2.0, kps[i].pt.y * 2.0, 2.0);
                const double d = mp.position.norm();
                mp.maxViewingDistance = (float)(d * std::pow(1.2, kps[i].octave + 0.5));
                mp.minViewingDistance = (float)(d * 0.1);
                mp.descriptor = kps[i].descriptor;
                db2.mapPoints.emplace(mp.id, mp);
                kf->mapPoints[i] = mp.id;
            }
        matchMapPointsSim3(*s1, *s2, Matrix4d::Identity(), db2, pairs, settings, c);
        if (pairs.size() < (size_t)next / 4) return 12;
        for (const auto &pr : pairs)
            if (db2.mapPoints.at(pr.first).observations.at(s1->id).v != db2.mapPoints.at(pr.second).observations.at(s2->id).v) {
                // a twin keypoint (same pixel row, identical descriptor) may legitimately win: distances must then be equal
                const auto &a = kps[(size_t)db2.mapPoints.at(pr.first).observations.at(s1->id).v], &b = kps[(size_t)db2.mapPoints.at(pr.second).observations.at(s2->id).v];
                if (a.descriptor != b.descriptor) return 13;
            }
        // triangulation matcher between the two views (identical bearings, translated camera): runs and returns pairs
        for (auto *kf : {s1.get(), s2.get()})
            for (size_t i = 0; i < kps.size(); ++i) {
                Vector3d ray;
                cam->pixelToRay(Vector2d(kps[i].pt.x, kps[i].pt.y), ray);
                kf->shared->keyPoints[i].bearing = ray;
                kf->shared->bowFeatureVec[kps[i].descriptor[0] % 5u].push_back((unsigned)i);
            }
        s2->poseCW(0, 3) = 0.3;
        const auto tri = matchForTriangulationDBoW(*s1, *s2, settings, c);
        for (const auto &pr : tri)
            if (s1->mapPoints[(size_t)pr.first.v].v != -1 || s2->mapPoints[(size_t)pr.second.v].v != -1) return 14;
        // batched MapPoint::updateDescriptor: a point seen once keeps that observation's descriptor
        std::vector<MpId> all;
        for (const auto &e : db2.mapPoints) all.push_back(e.first);
        updateDescriptors(db2, all, c);
        for (const auto &e : db2.mapPoints) {
            const auto &obs = *e.second.observations.begin();
            if (e.second.descriptor != db2.keyframes.at(obs.first)->shared->keyPoints[(size_t)obs.second.v].descriptor) return 15;
        }
        std::printf("matchers ok: %d projected, %u fused, %zu sim3 pairs, %zu triangulation pairs\n", found, fused, pairs.size(), tri.size());
    }
    // ---- DBoW2 text vocabulary loader (bow_index.cpp:11-19) on the file the test wrote -------------------------------
    {
        BowVocabulary tv;
        if (loadVocabularyText(out + "/voc.txt", tv)) {
            // (the text format does not store the root's descriptor: compare from node 1 on)
            if (tv.childOff != voc.childOff || tv.childIds != voc.childIds || tv.nodeWord != voc.nodeWord
                || tv.nodeDescriptor.size() != voc.nodeDescriptor.size()
                || !std::equal(tv.nodeDescriptor.begin() + 8, tv.nodeDescriptor.end(), voc.nodeDescriptor.begin() + 8)
                || tv.nodeWeight != voc.nodeWeight || tv.levels != voc.levels) return 16;
            std::printf("vocabulary text loader ok: %zu nodes\n", tv.nodeWord.size());
        } else if (!voc.childOff.empty()) return 17;
    }
    std::printf("ok %zu %zu %u %u\n", kpsA.size(), kpsB.size(), n, nb);
    return 0;
}
