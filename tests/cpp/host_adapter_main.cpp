// Drives the C++ adapters (slam-module_b200/host/slam_frontend.hpp) the way the reference's callers drive
// the original classes (keyframe.cpp:95-116, mapper_helpers.cpp:1190-1194, loop_closer.cpp:195) and dumps
// every artefact as raw binary; tests/test_gpu_host_adapters.py compares the dumps with the CPU oracle.
//
//   host_adapter_main <w> <h> <imgA.raw> <imgB.raw> <outdir> <maxKeypoints>
#include <cstdio>
#include <cstdlib>
#include <fstream>
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
        if (id.v >= 0) { MapPoint mp; mp.status = (i % 7 == 6) ? MapPointStatus::NOT_TRIANGULATED : MapPointStatus::TRIANGULATED; db1.mapPoints.push_back(mp); }
        kf1.mapPoints.push_back(id);
    }
    for (size_t i = 0; i < kf2.shared->keyPoints.size(); ++i) {
        MpId id; id.v = (int)db2.mapPoints.size();
        MapPoint mp; mp.status = (i % 3 == 2) ? MapPointStatus::NOT_TRIANGULATED : MapPointStatus::TRIANGULATED;
        db2.mapPoints.push_back(mp);
        kf2.mapPoints.push_back(id);
    }
    auto fe = cudaFrontend(settings, w, h);
    std::vector<int> matched;
    const unsigned n = matchForLoopClosures(kf1, kf2, db1, db2, matched, params.slam, cudaContext(fe));
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
    std::printf("ok %zu %zu %u %u\n", kpsA.size(), kpsB.size(), n, nb);
    return 0;
}
