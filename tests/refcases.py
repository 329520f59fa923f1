"""Seeded scenarios that run BOTH the oracle restatement (oracle/src/*.cpp) and the reference's own sources compiled
verbatim (oracle/_ref/libref_slam.so, oracle/ref_slam.cpp) on the same inputs.

Every case is a function `case_xxx(backend)` with backend in {"oracle", "ref"} returning a dict of numpy arrays; the
two dicts must be equal key by key.  tools/gen_golden.py stores the "ref" dicts in tests/golden/golden_ref_slam.npz so
that the comparison also runs where the reference tree (and the prebuilt oracle/_ref) is absent.
"""
import numpy as np

import slam_module_b200 as sm
from oracle import pyoracle as po
from oracle import pyref as pr


class OracleImpl:
    """The CPU oracle behind the method names the cases call."""
    name = "oracle"
    match_bow = staticmethod(po.match_bow)
    match_bruteforce = staticmethod(po.match_bruteforce)
    match_triangulation = staticmethod(po.match_triangulation)
    search_candidates = staticmethod(po.search_candidates)
    match_sim3 = staticmethod(po.match_sim3)
    medoid = staticmethod(po.medoid)

    def geometry(self, lv, sf, mk):
        s, _, _, b = po.geometry(po.make_params(640, 480, levels=lv, scale_factor=sf, max_keypoints=mk))
        return s, b

    def feature_order(self, x, y):
        return po.feature_index(x, y)

    def extract(self, p, img, tracks, ids, level):
        return po.extract(p, img, tracks, ids, level)

    def pyramid(self, p, img):
        return po.pyramid(p, img)

    def bow(self, vocab, desc, kfs, removed, queries, n_words):
        word, weight, node = po.bow_transform(vocab, desc, levels_up=4)
        w, v = po.bow_vector(word, weight)
        ix = po.BowIndex(n_words)
        for k, (ww, vv) in enumerate(kfs):
            ix.add(1000, k, ww, vv)
        for k in removed:
            ix.remove(1000, k)
        sims = [ix.similar(kfs[k][0], kfs[k][1], self_key=(1000, k))[1:] for k in queries]
        ix.close()
        return word, weight, node, w, v, sims


class GpuImpl:
    """libslamgpu.so through its C ABI (slam_module_b200.slamgpu): the product under test."""
    name = "gpu"

    def __init__(self, slamgpu):
        self.sg = slamgpu
        self.ctx = slamgpu.Context(640, 480, max_frames=1)
        for m in ("match_bow", "match_bruteforce", "match_triangulation", "search_candidates", "match_sim3", "medoid"):
            setattr(self, m, getattr(self.ctx, m))

    def close(self):
        self.ctx.close()

    def geometry(self, lv, sf, mk):
        side = max(64, int(48 * sf ** (lv - 1)) + 1)      # the library refuses pyramids whose top level cannot hold a patch
        with self.sg.Context(side, side, levels=lv, scale_factor=sf, max_keypoints=mk, max_frames=1) as c:
            return c.scales.copy(), c.budgets.copy()

    def feature_order(self, x, y):
        return self.sg.feature_index(x, y)

    def _ctx(self, p, max_tracks=0, level=0):
        return self.sg.Context(p.width, p.height, levels=p.levels, scale_factor=p.scale_factor, max_keypoints=p.max_keypoints,
                               ini_fast_thr=p.ini_fast_thr, min_fast_thr=p.min_fast_thr, max_frames=1, max_tracks=max_tracks,
                               track_level=level)

    def extract(self, p, img, tracks, ids, level):
        nt = 0 if tracks is None else len(tracks)
        with self._ctx(p, nt, level) as c:
            return c.detect_and_extract(img, None if tracks is None else [tracks], None if ids is None else [ids])[0]

    def pyramid(self, p, img):
        with self._ctx(p) as c:
            c.pyramid_update(img)
            return ([c.get_level(0, l) for l in range(p.levels)], [c.get_blurred_level(0, l) for l in range(p.levels)])

    def bow(self, vocab, desc, kfs, removed, queries, n_words):
        voc = self.sg.Vocabulary(self.ctx, vocab)
        word, weight, node = voc.transform(desc, levels_up=4)
        w, v = voc.bow_vector(word, weight)
        db = self.sg.BowDatabase(self.ctx, len(kfs) + 4)
        for k, (ww, vv) in enumerate(kfs):
            db.add(1000, k, ww, vv)
        for k in removed:
            db.remove(1000, k)
        sims = [db.similar(kfs[k][0], kfs[k][1], self_key=(1000, k))[1:] for k in queries]
        db.close(); voc.close()
        return word, weight, node, w, v, sims


def _impl(backend):
    return OracleImpl() if backend == "oracle" else backend


def _flip(rng, d, nbits):
    d = d.copy()
    for b in rng.integers(0, 256, nbits):
        d[b >> 5] ^= np.uint32(1 << (int(b) & 31))
    return d


# ---- C1 / C2: StaticSettings -------------------------------------------------------------------------------------
SETTINGS = [(8, 1.2, 1000), (8, 1.2, 2000), (4, 1.5, 500), (12, 1.1, 3000), (1, 1.2, 100), (8, 2.0, 800), (6, 1.3, 7)]


def case_settings(backend):
    out = {}
    for i, (lv, sf, mk) in enumerate(SETTINGS):
        if backend == "ref":
            s, _, b = pr.settings(lv, sf, mk)
        else:
            s, b = _impl(backend).geometry(lv, sf, mk)
        out["scale%d" % i] = s
        out["budget%d" % i] = b
    return out


# ---- F1: FeatureSearch -------------------------------------------------------------------------------------------
def case_feature_search(backend):
    rng = np.random.default_rng(21)
    n = 700
    x = rng.uniform(0, 640, n).astype(np.float32)
    y = np.round(rng.uniform(0, 480, n)).astype(np.float32)       # integer y: many ties in the Y sort
    x[::9] = np.round(x[::9])
    if backend not in ("oracle", "ref"):     # the C ABI exposes the index order (queries: sg_search_candidates cases below)
        return {"order": backend.feature_order(x, y)}
    f = po.features_around if backend == "oracle" else pr.features_around
    out = {"order": f(x, y, 320.0, 240.0, 2000.0)}                 # huge radius: the whole Y-sorted order
    qs = [(10.0, 10.0, 30.0), (320.0, 240.0, 0.0), (320.5, 100.0, 1.0), (600.0, 470.0, 55.5), (-20.0, 200.0, 40.0),
          (x[5], y[5], 7.0), (x[77], y[77], 12.25)]
    for i, (qx, qy, r) in enumerate(qs):
        out["q%d" % i] = f(x, y, qx, qy, r)
    return out


# ---- P1 / P2 / D1 / O1-O5: pyramid and OrbExtractor::detectAndExtract ---------------------------------------------
EXTRACT = {
    "vga2000": dict(w=640, h=480, seed=1000, kw=dict(max_keypoints=2000)),
    "vga1000": dict(w=640, h=480, seed=1001, kw=dict(max_keypoints=1000)),
    "odd": dict(w=333, h=251, seed=7, kw=dict(max_keypoints=600, levels=5, scale_factor=1.3)),
    "tracks": dict(w=640, h=480, seed=1002, kw=dict(max_keypoints=500), tracks=True),
}


def _extract_inputs(name):
    c = EXTRACT[name]
    p = po.make_params(c["w"], c["h"], **c["kw"])
    img = sm.synth.frame(c["w"], c["h"], c["seed"])
    tracks = ids = None
    level = 0
    if c.get("tracks"):
        rng = np.random.default_rng(5)
        tracks = np.stack([rng.uniform(-5, c["w"] + 5, 160), rng.uniform(-5, c["h"] + 5, 160)], axis=1).astype(np.float32)
        tracks[::7] = np.round(tracks[::7]) + 0.5                 # x.5: cvRound (half-even) vs the border filter
        ids = rng.permutation(1000)[:160].astype(np.int32)
        level = 1
    return p, img, tracks, ids, level


def case_extract(backend, name):
    p, img, tracks, ids, level = _extract_inputs(name)
    r = pr.extract(p, img, tracks, ids, level) if backend == "ref" else _impl(backend).extract(p, img, tracks, ids, level)
    return {k: r[k] for k in ("x", "y", "angle", "octave", "desc", "track_id")}


def case_pyramid_crc(backend):
    import zlib
    p = po.make_params(640, 480)
    img = sm.synth.frame(640, 480, 1000)
    lv, bl = pr.pyramid(p, img) if backend == "ref" else _impl(backend).pyramid(p, img)
    return {"pyr": np.array([zlib.crc32(a.tobytes()) for a in lv], np.int64),
            "blur": np.array([zlib.crc32(a.tobytes()) for a in bl], np.int64)}


# ---- M1: matchForLoopClosures --------------------------------------------------------------------------------------
def _bow_sets(seed, n=900, n_nodes=40):
    rng = np.random.default_rng(seed)
    dA = rng.integers(0, 2 ** 32, (n, 8), dtype=np.uint32)
    aA = rng.uniform(0, 360, n).astype(np.float32)
    src = rng.permutation(n)
    dB = np.stack([_flip(rng, dA[i], int(rng.integers(0, 28))) if rng.random() < 0.75 else rng.integers(0, 2 ** 32, 8, dtype=np.uint32)
                   for i in src])
    aB = ((aA[src] - 30 + rng.normal(0, 4, n)) % 360).astype(np.float32)
    aB[rng.random(n) < 0.15] = rng.uniform(0, 360)                 # wrong-angle matches: the histogram rejects them
    # near-identical descriptor clusters: second-best ties and the uniqueness rule
    dB[5::11] = dB[4::11][:len(dB[5::11])]
    nodeA = rng.integers(0, n_nodes, n).astype(np.int32) * 3
    # a copy mostly lands under its source's node
    nodeB = np.where(rng.random(n) < 0.85, nodeA[src], rng.integers(0, n_nodes, n) * 3).astype(np.int32)
    nodeB[5::11] = nodeB[4::11][:len(nodeB[5::11])]
    nodeA[rng.random(n) < 0.03] = -1
    nodeB[rng.random(n) < 0.03] = -1
    nodeB[nodeB == 30] = 31                                        # a node only one side has: lower_bound skips
    sA = rng.choice([0, 1, 2], n, p=[0.1, 0.75, 0.15]).astype(np.uint8)
    sB = rng.choice([0, 1, 2], n, p=[0.1, 0.8, 0.1]).astype(np.uint8)
    return dA, aA, nodeA, sA, dB, aB, nodeB, sB


def case_loop_closures(backend, seed, require):
    dA, aA, nodeA, sA, dB, aB, nodeB, sB = _bow_sets(seed)
    if backend == "ref":
        n, m = pr.match_loop_closures(dA, aA, nodeA, dB, aB, nodeB, sA, sB, ratio=0.8, require_triangulation=require)
    else:
        eA = ((sA == 1) | ((sA == 2) & (not require))).astype(np.uint8)       # keyframe_matcher.cpp:79-84
        eB = (sB == 1).astype(np.uint8)                                        # :94-96
        n, m = _impl(backend).match_bow(dA, aA, nodeA, dB, aB, nodeB, eA, eB, ratio=0.8, thr=50, check_orientation=True)
    return {"n": np.array([n]), "matches": m}


def case_loop_closures_bruteforce(backend, seed):
    """Single node holding every feature, every feature eligible: the brute-force degenerate case (BASELINE config 3)."""
    dA, aA, dB, aB = sm.synth.correlated_descriptors(2000, seed)
    zA = np.zeros(len(dA), np.int32)
    zB = np.zeros(len(dB), np.int32)
    if backend == "ref":
        n, m = pr.match_loop_closures(dA, aA, zA, dB, aB, zB)
    else:
        n, m = _impl(backend).match_bruteforce(dA, aA, dB, aB, ratio=0.8, thr=50, check_orientation=True)
    return {"n": np.array([n]), "matches": m}


# ---- M2: matchForTriangulationDBoW -----------------------------------------------------------------------------------
def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _triangulation_scene(seed, n=700, n_nodes=25):
    rng = np.random.default_rng(seed)
    P = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 12, n)], axis=1)
    poseA = np.eye(4)
    poseB = np.eye(4)
    poseB[:3, :3] = _rot(0.02, -0.05, 0.01)
    poseB[:3, 3] = [0.6, 0.05, -0.1]
    bA = P / np.linalg.norm(P, axis=1, keepdims=True)
    PB = P @ poseB[:3, :3].T + poseB[:3, 3]
    PB = PB + rng.normal(0, 0.004, PB.shape) * PB[:, 2:3] * (rng.random((n, 1)) < 0.5)   # half the rays are noisy
    bB = PB / np.linalg.norm(PB, axis=1, keepdims=True)
    perm = rng.permutation(n)
    dA = rng.integers(0, 2 ** 32, (n, 8), dtype=np.uint32)
    dB = np.stack([_flip(rng, dA[i], int(rng.integers(0, 40))) for i in perm])
    bB = bB[perm]
    dB[3::13] = dB[2::13][:len(dB[3::13])]                          # equal-distance candidates: the '<=' keeps the LAST one
    aA = rng.uniform(0, 360, n).astype(np.float32)
    aB = ((aA[perm] - 20 + rng.normal(0, 5, n)) % 360).astype(np.float32)
    octA = rng.integers(0, 8, n).astype(np.int32)
    nodeA = rng.integers(0, n_nodes, n).astype(np.int32)
    nodeB = np.where(rng.random(n) < 0.9, nodeA[perm], rng.integers(0, n_nodes, n)).astype(np.int32)
    hA = (rng.random(n) < 0.2).astype(np.uint8)
    hB = (rng.random(n) < 0.2).astype(np.uint8)
    return dA, aA, octA, bA, nodeA, hA, dB, aB, bB, nodeB, hB, poseA, poseB


def case_triangulation(backend, seed, thr_deg):
    dA, aA, octA, bA, nodeA, hA, dB, aB, bB, nodeB, hB, poseA, poseB = _triangulation_scene(seed)
    # the essential matrix is an input of the oracle (and of the C ABI): take the one the reference builds from the poses
    # when it is available, else the same formula in numpy (essential_solver.cc:157-162)
    if backend == "ref":
        n, m, E = pr.match_triangulation(dA, aA, octA, bA, nodeA, dB, aB, bB, nodeB, poseA, poseB, hA, hB, residual_deg_thr=thr_deg)
        return {"n": np.array([n]), "matches": m, "E": E}
    E = essential_from_poses(poseA, poseB)
    sf = po.geometry(po.make_params(640, 480))[0]
    n, m = _impl(backend).match_triangulation(dA, aA, octA, bA, nodeA, dB, aB, bB, nodeB, E, sf, (1 - hA).astype(np.uint8),
                                              (1 - hB).astype(np.uint8), residual_deg_thr=thr_deg)
    return {"n": np.array([n]), "matches": m, "E": E}


def essential_from_poses(pose1, pose2):
    """create_E_21(rot_2w, trans_2w, rot_1w, trans_1w) as matchForTriangulationDBoW calls it (keyframe_matcher.cpp:171-175),
    with Eigen's x0 + (x1 + x2) summation of every 3-term inner product."""
    def s3(a, b):      # row . column with the tree order
        p = a * b
        return p[0] + (p[1] + p[2])

    def mm(A, B):
        return np.array([[s3(A[r], B[:, c]) for c in range(B.shape[1])] for r in range(3)])
    r1, t1 = pose2[:3, :3], pose2[:3, 3:4]      # "1" of create_E_21 is keyframe 2
    r2, t2 = pose1[:3, :3], pose1[:3, 3:4]
    rot21 = mm(r2, r1.T)
    t21 = mm(-rot21, t1) + t2
    t = t21[:, 0]
    zero = 0.0
    skew = np.array([[zero, -t[2], t[1]], [t[2], zero, -t[0]], [-t[1], t[0], zero]])
    return mm(skew, rot21)


# ---- M3 / M4: projection matchers -------------------------------------------------------------------------------------
def _projection_scene(seed, nk=800, nq=500, cluster=True):
    """Keypoints + map points whose projection (identity pose, f = 1 pinhole: pixel = position.xy / position.z with
    z = 1) lands near a keypoint.  Returns keypoint arrays and the map-point arrays of oracle/ref_slam.cpp."""
    rng = np.random.default_rng(seed)
    kx = rng.uniform(20, 620, nk).astype(np.float32)
    ky = (np.round(rng.uniform(20, 460, nk) * 2) / 2).astype(np.float32)
    koct = rng.integers(0, 8, nk).astype(np.int32)
    kdesc = rng.integers(0, 2 ** 32, (nk, 8), dtype=np.uint32)
    if cluster:
        kx[1::7] = kx[0::7][:len(kx[1::7])] + 0.5
        ky[1::7] = ky[0::7][:len(ky[1::7])]
        kdesc[1::7] = kdesc[0::7][:len(kdesc[1::7])]
        koct[1::7] = koct[0::7][:len(koct[1::7])]
    src = rng.integers(0, nk, nq)
    qdesc = np.stack([_flip(rng, kdesc[s], int(rng.integers(0, 70))) for s in src])
    px = (kx[src] + rng.normal(0, 2.5, nq)).astype(np.float32).astype(np.float64)
    py = (ky[src] + rng.normal(0, 2.5, nq)).astype(np.float32).astype(np.float64)
    far = rng.random(nq) < 0.08
    px[far] = rng.uniform(0, 640, far.sum())
    pos = np.stack([px, py, np.ones(nq)], axis=1)
    behind = rng.random(nq) < 0.03
    pos[behind, 2] = -1.0                                           # behind the camera: rejected by reprojection
    dist = np.linalg.norm(pos, axis=1)
    level = np.clip(koct[src] + rng.integers(-1, 2, nq), 0, 7)
    # maxViewingDistance = dist * 1.2^(level - 0.5): predictScaleLevel = ceil(log(max / dist) / log 1.2) = level
    max_d = (dist * 1.2 ** (level - 0.5)).astype(np.float32)
    max_d[level == 0] = (dist[level == 0] * 0.99).astype(np.float32)   # ratio < 1 -> clamped to level 0 ... but out of range
    max_d[level == 0] = (dist[level == 0] * 1.0001).astype(np.float32)
    min_d = (dist * 0.2).astype(np.float32)
    out_of_range = rng.random(nq) < 0.04
    min_d[out_of_range] = (dist[out_of_range] * 1.5).astype(np.float32)
    view = -pos / dist[:, None]
    # tilt a third of the normals by ~10 degrees (cos < 0.998: the full search radius) and a few beyond 60 degrees
    tilt = rng.random(nq)
    ang = np.where(tilt < 0.33, np.deg2rad(10.0), np.where(tilt < 0.38, np.deg2rad(70.0), 0.0))
    axis = np.cross(view, np.array([0.0, 0.0, 1.0]) + 1e-3)
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    norm = view * np.cos(ang)[:, None] + np.cross(axis, view) * np.sin(ang)[:, None]
    zero_norm = rng.random(nq) < 0.03
    norm[zero_norm] = 0.0
    return kx, ky, koct, kdesc, pos, norm.astype(np.float32), min_d, max_d, qdesc


def case_search_by_projection(backend, seed, golden=None):
    kx, ky, koct, kdesc, pos, norm, min_d, max_d, qdesc = _projection_scene(seed)
    taken = (np.arange(len(kx)) % 5 == 0).astype(np.uint8)
    if backend == "ref":
        n, idx, qx, qy, qr, ql = pr.search_by_projection(kx, ky, koct, kdesc, pos, norm, min_d, max_d, qdesc, 15.0, taken)
        return {"n": np.array([n]), "idx": idx, "qx": qx, "qy": qy, "qr": qr, "ql": ql}
    # the oracle takes the projected queries: those the reference computed (golden: stored with the fixture)
    g = golden
    live = g["qr"] >= 0
    t = taken.copy()
    idx = np.full(len(qdesc), -1, np.int32)
    n, i2, _ = _impl(backend).search_candidates(kx, ky, koct, kdesc, g["qx"][live], g["qy"][live], g["qr"][live], qdesc[live],
                                                mode=1, thr=100, taken=t)
    idx[live] = i2
    return {"n": np.array([n]), "idx": idx, "qx": g["qx"], "qy": g["qy"], "qr": g["qr"], "ql": g["ql"]}


def case_replace_duplication(backend, seed, golden=None):
    """Keypoints without map points and queries that each pick a different keypoint cluster member at most once is NOT
    guaranteed: the fuse bookkeeping of :500-524 is simulated from the oracle's best-candidate list."""
    kx, ky, koct, kdesc, pos, norm, min_d, max_d, qdesc = _projection_scene(seed, nk=900, nq=400)
    rng = np.random.default_rng(seed + 1)
    kp_mp = np.where(rng.random(len(kx)) < 0.3, rng.integers(1, 4, len(kx)), 0).astype(np.int32)
    q_obs = rng.integers(0, 4, len(qdesc)).astype(np.int32)
    if backend == "ref":
        n, fin, qx, qy, qr, ql = pr.replace_duplication(kx, ky, koct, kdesc, pos, norm, min_d, max_d, qdesc, 3.0, kp_mp, q_obs)
        return {"n": np.array([n]), "final": fin, "qx": qx, "qy": qy, "qr": qr, "ql": ql}
    g = golden
    live = g["qr"] >= 0
    best = np.full(len(qdesc), -1, np.int32)
    _, b2, _ = _impl(backend).search_candidates(kx, ky, koct, kdesc, g["qx"][live], g["qy"][live], g["qr"][live], qdesc[live],
                                                mode=0, thr=50)
    best[live] = b2
    # keyframe_matcher.cpp:500-524 on observation counts: owner[k] >= 0 query index, -2 original map point, -1 none
    owner = np.where(kp_mp > 0, -2, -1).astype(np.int64)
    n_obs_kp = kp_mp.astype(np.int64).copy()          # observations of the point that owns keypoint k
    fused = 0
    for q in range(len(qdesc)):
        k = int(best[q])
        if k < 0:
            continue
        if owner[k] == -1:
            owner[k] = q
            n_obs_kp[k] = q_obs[q] + 1
        else:
            mine, theirs = int(q_obs[q]), int(n_obs_kp[k])
            if mine < theirs:
                pass                                   # the query point is replaced by the keypoint's: owner unchanged
                n_obs_kp[k] = theirs + mine            # replaceWith moves the query's other observations over
            else:
                owner[k] = q                           # the keypoint's point is replaced by the query point
                n_obs_kp[k] = mine + theirs
        fused += 1
    return {"n": np.array([fused]), "final": owner.astype(np.int32), "qx": g["qx"], "qy": g["qy"], "qr": g["qr"], "ql": g["ql"]}


def _sim3_scene(seed, n=600):
    rng = np.random.default_rng(seed)
    # physical points seen by both keyframes at slightly different pixels
    X = rng.uniform(30, 610, n)
    Y = np.round(rng.uniform(30, 450, n) * 2) / 2
    base = rng.integers(0, 2 ** 32, (n, 8), dtype=np.uint32)
    oct0 = rng.integers(0, 8, n)
    x1 = (X + rng.normal(0, 1.0, n)).astype(np.float32); y1 = (Y + rng.normal(0, 1.0, n)).astype(np.float32)
    p2 = rng.permutation(n)
    x2 = (X[p2] + rng.normal(0, 1.0, n)).astype(np.float32); y2 = (Y[p2] + rng.normal(0, 1.0, n)).astype(np.float32)
    d1 = np.stack([_flip(rng, base[i], int(rng.integers(0, 25))) for i in range(n)])
    d2 = np.stack([_flip(rng, base[i], int(rng.integers(0, 25))) for i in p2])
    oct1 = np.clip(oct0 + rng.integers(-1, 1, n), 0, 7).astype(np.int32)
    oct2 = np.clip(oct0[p2] + rng.integers(-1, 1, n), 0, 7).astype(np.int32)
    # map points: one per keypoint of kf1 (ids 0..n-1) and of kf2 (ids n..2n-1), 15 % of the keypoints own none
    mp1 = np.where(rng.random(n) < 0.85, np.arange(n), -1).astype(np.int32)
    mp2 = np.where(rng.random(n) < 0.85, n + np.arange(n), -1).astype(np.int32)
    pos = np.concatenate([np.stack([x1.astype(np.float64), y1.astype(np.float64), np.ones(n)], axis=1),
                          np.stack([x2.astype(np.float64), y2.astype(np.float64), np.ones(n)], axis=1)])
    pos[:, :2] += rng.normal(0, 1.5, (2 * n, 2))
    pos[rng.random(2 * n) < 0.02, 2] = -1.0
    dist = np.linalg.norm(pos, axis=1)
    lvl = np.concatenate([oct1, oct2]) + rng.integers(0, 2, 2 * n)
    lvl = np.clip(lvl, 0, 7)
    mx = (dist * 1.2 ** (lvl - 0.5)).astype(np.float32)
    mx[lvl == 0] = (dist[lvl == 0] * 1.0001).astype(np.float32)
    mn = (dist * 0.2).astype(np.float32)
    mn[rng.random(2 * n) < 0.03] *= 10
    desc = np.concatenate([np.stack([_flip(rng, d1[i], int(rng.integers(0, 10))) for i in range(n)]),
                           np.stack([_flip(rng, d2[i], int(rng.integers(0, 10))) for i in range(n)])])
    status = rng.choice([0, 1], 2 * n, p=[0.9, 0.1]).astype(np.int32)      # MapPointStatus: 0 TRIANGULATED, 1 NOT_TRIANGULATED
    # seeds: a few already matched (kf1 point, kf2 point) pairs of the same physical point
    inv = np.empty(n, np.int64); inv[p2] = np.arange(n)
    seeds = [(int(mp1[i]), int(mp2[inv[i]])) for i in range(0, n, 17) if mp1[i] >= 0 and mp2[inv[i]] >= 0]
    return x1, y1, oct1, d1, mp1, x2, y2, oct2, d2, mp2, pos, mn, mx, desc, status, np.array(seeds, np.int32).reshape(-1, 2)


def case_sim3(backend, seed, golden=None):
    x1, y1, oct1, d1, mp1, x2, y2, oct2, d2, mp2, pos, mn, mx, desc, status, seeds = _sim3_scene(seed)
    if backend == "ref":
        pairs, q12, l12, q21, l21 = pr.match_sim3(x1, y1, oct1, d1, mp1, x2, y2, oct2, d2, mp2, pos, mn, mx, desc, status, seeds)
        return {"pairs": pairs, "q12": q12, "l12": l12, "q21": q21, "l21": l21}
    g = golden
    qd12 = desc[np.maximum(mp1, 0)]
    qd21 = desc[np.maximum(mp2, 0)]
    pairs = _impl(backend).match_sim3(x1, y1, oct1, d1, x2, y2, oct2, d2, g["q12"], qd12, g["l12"], g["q21"], qd21, g["l21"])
    return {"pairs": pairs, "q12": g["q12"], "l12": g["l12"], "q21": g["q21"], "l21": g["l21"]}


# ---- f2: MapPoint::updateDescriptor ---------------------------------------------------------------------------------
def case_medoid(backend):
    rng = np.random.default_rng(31)
    sizes = [1, 2, 3, 4, 7, 12, 33, 5, 2, 64]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    base = rng.integers(0, 2 ** 32, 8, dtype=np.uint32)
    desc = np.stack([_flip(rng, base, int(rng.integers(0, 90))) for _ in range(int(offs[-1]))])
    desc[offs[5] + 3] = desc[offs[5] + 1]                           # twins: first index wins
    if backend == "ref":
        return {"desc": pr.medoid(desc, offs)}
    best = _impl(backend).medoid(desc, offs)
    return {"desc": np.stack([desc[offs[s] + best[s]] for s in range(len(sizes))])}


# ---- f3: BowIndex over a DBoW2 text vocabulary -----------------------------------------------------------------------
def write_vocabulary_txt(vocab, path, branching):
    """DBoW2 / ORB-SLAM text format: 'k L scoring weighting' then one line per non-root node in id order:
    parent is_leaf d0 .. d31 weight.  synth.random_vocabulary stores nodes breadth first with contiguous ids."""
    n = len(vocab["node_word"])
    parent = np.zeros(n, np.int64)
    for i in range(n):
        for c in vocab["child_ids"][vocab["child_off"][i]:vocab["child_off"][i + 1]]:
            parent[c] = i
    with open(path, "w") as f:
        f.write("%d %d 0 0\n" % (branching, vocab["levels"]))
        for i in range(1, n):
            by = vocab["node_desc"][i].view(np.uint8)
            f.write("%d %d %s %r\n" % (parent[i], 1 if vocab["node_word"][i] >= 0 else 0, " ".join(str(int(b)) for b in by),
                                        float(vocab["node_weight"][i])))


def case_bow(backend, tmpdir):
    vocab = sm.synth.random_vocabulary(6, 4, 41)
    rng = np.random.default_rng(43)
    leaves = np.flatnonzero(vocab["node_word"] >= 0)
    desc = np.stack([_flip(rng, vocab["node_desc"][leaves[int(rng.integers(0, len(leaves)))]], int(rng.integers(0, 30)))
                     for _ in range(1500)])
    n_words = int((vocab["node_word"] >= 0).sum())
    kfs = sm.synth.random_bow_vectors(120, n_words, 60, 47)
    out = {}
    if backend == "ref":
        path = str(tmpdir / "vocab.txt")
        write_vocabulary_txt(vocab, path, 6)
        ix = pr.BowIndex(path)
        node, w, v = ix.transform(desc)
        out.update(node=node, word=w, value=v)
        for k, (ww, vv) in enumerate(kfs):
            ix.add(k, ww, vv)
        for k in (3, 50, 51):
            ix.remove(k)
        for qi, k in enumerate((0, 7, 99, 3)):
            kf, sc = ix.similar(kfs[k][0], kfs[k][1], self_kf=k)
            out["sim%d_kf" % qi] = kf
            out["sim%d_score" % qi] = sc
        ix.close()
        return out
    word, weight, node, w, v, sims = _impl(backend).bow(vocab, desc, kfs, (3, 50, 51), (0, 7, 99, 3), n_words)
    # DBoW2 drops weight-0 features from the feature vector; the reference tree is numbered by its text file (node id ==
    # line order == breadth-first id of synth.random_vocabulary), so node ids compare directly
    out.update(node=np.where(weight > 0, node, -1).astype(np.int32), word=w, value=v)
    for qi, (kf, sc) in enumerate(sims):
        out["sim%d_kf" % qi] = kf
        out["sim%d_score" % qi] = sc
    return out
