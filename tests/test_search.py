"""Candidate-list matchers and descriptor medoid: the oracle against plain-python restatements of
the reference loops (CPU), and the CUDA library against the oracle (GPU, through the C ABI)."""
import numpy as np
import pytest


def _popc(a, b):
    return int(sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(a, b)))


def _scene(seed, nk=600, nq=300, w=640, h=480):
    """Keypoints of one keyframe and projected map points: queries are noisy copies of keypoint descriptors
    placed near their keypoint (so matches exist), some far away, some with duplicated y (sort ties)."""
    rng = np.random.default_rng(seed)
    kx = rng.uniform(0, w, nk).astype(np.float32)
    ky = np.round(rng.uniform(0, h, nk) * 2) / 2            # half-pixel grid: many equal y values
    ky = ky.astype(np.float32)
    koct = rng.integers(0, 8, nk).astype(np.int32)
    kdesc = rng.integers(0, 2 ** 32, (nk, 8), dtype=np.uint32)
    src = rng.integers(0, nk, nq)
    qdesc = kdesc[src].copy()
    for q in range(nq):
        for b in rng.integers(0, 256, int(rng.integers(0, 60))):
            qdesc[q, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    qx = (kx[src] + rng.normal(0, 3, nq)).astype(np.float32)
    qy = (ky[src] + rng.normal(0, 3, nq)).astype(np.float32)
    far = rng.random(nq) < 0.1
    qx[far] = rng.uniform(-50, w + 50, far.sum()).astype(np.float32)
    qr = rng.uniform(2, 25, nq).astype(np.float32)
    qlevel = np.clip(koct[src] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32)
    # near-duplicate keypoints: several candidates with the same distance inside one radius
    kx[1::7] = kx[0::7][:len(kx[1::7])] + 0.5
    ky[1::7] = ky[0::7][:len(ky[1::7])]
    kdesc[1::7] = kdesc[0::7][:len(kdesc[1::7])]
    return kx, ky, koct, kdesc, qx, qy, qr, qdesc, qlevel


def test_oracle_feature_search_matches_python(oracle):
    kx, ky, koct, kdesc, qx, qy, qr, qdesc, _ = _scene(1, nk=300, nq=40)
    order = oracle.feature_index(kx, ky)
    assert sorted(order.tolist()) == list(range(300))
    assert (np.diff(ky[order]) >= 0).all()
    for q in range(40):
        got = oracle.features_around(kx, ky, qx[q], qy[q], qr[q])
        want = [int(i) for i in order if ky[i] >= np.float32(qy[q] - qr[q]) and ky[i] <= np.float32(qy[q] + qr[q])
                and np.float32(np.float32((qx[q] - kx[i]) * (qx[q] - kx[i])) + np.float32((qy[q] - ky[i]) * (qy[q] - ky[i])))
                < np.float32(qr[q] * qr[q])]
        assert got.tolist() == want


def test_oracle_search_and_medoid_match_python(oracle):
    kx, ky, koct, kdesc, qx, qy, qr, qdesc, qlevel = _scene(2, nk=200, nq=60)
    # mode 1 (searchByProjection), plain python
    taken = (np.arange(200) % 5 == 0).astype(np.uint8)
    tk = taken.copy()
    want_idx = []
    for q in range(60):
        cand = oracle.features_around(kx, ky, qx[q], qy[q], qr[q])
        best, best2, lvl, lvl2, bi = 256, 256, -1, -1, -1
        for i in cand:
            if tk[i]:
                continue
            d = _popc(qdesc[q], kdesc[i])
            if d < best:
                best2, best, lvl2, lvl, bi = best, d, lvl, int(koct[i]), int(i)
            elif d < best2:
                lvl2, best2 = int(koct[i]), d
        ok = bi != -1 and best <= 100 and not (lvl == lvl2 and best > 0.8 * best2)
        want_idx.append(bi if ok else -1)
        if ok:
            tk[bi] = 1
    t2 = taken.copy()
    n, idx, dist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t2)
    assert idx.tolist() == want_idx and n == sum(i >= 0 for i in want_idx) and np.array_equal(t2, tk)
    assert n > 10
    # medoid, plain python
    rng = np.random.default_rng(3)
    sizes = [1, 2, 3, 7, 12, 0, 5]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    desc = rng.integers(0, 2 ** 32, (offs[-1], 8), dtype=np.uint32)
    want = []
    for s, n_ in enumerate(sizes):
        d = desc[offs[s]:offs[s + 1]]
        best_m, best_i = 256, 0
        for i in range(n_):
            row = sorted(_popc(d[i], d[j]) for j in range(n_))
            m = row[int(0.5 * (n_ - 1))]
            if m < best_m:
                best_m, best_i = m, i
        want.append(best_i)
    assert oracle.medoid(desc, offs).tolist() == want


@pytest.mark.gpu
@pytest.mark.parametrize("seed,nk,nq", [(11, 600, 300), (12, 2000, 1500), (13, 50, 400)])
def test_gpu_search_candidates_bit_exact(slamgpu, oracle, seed, nk, nq):
    kx, ky, koct, kdesc, qx, qy, qr, qdesc, qlevel = _scene(seed, nk, nq)
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        assert np.array_equal(slamgpu.feature_index(kx, ky), oracle.feature_index(kx, ky))
        # replaceDuplication: best only, threshold 50
        n, idx, dist = ctx.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=50)
        rn, ridx, rdist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=50)
        assert n == rn and np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and n > 0
        # findMatchesTranformedMps: threshold 100 + octave window
        n, idx, dist = ctx.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=100, qlevel=qlevel)
        rn, ridx, rdist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=0, thr=100, qlevel=qlevel)
        assert n == rn and np.array_equal(idx, ridx) and np.array_equal(dist, rdist)
        # searchByProjection: consumed keypoints, level-aware ratio rule, queries in order
        taken = (np.arange(nk) % 4 == 0).astype(np.uint8)
        t1, t2 = taken.copy(), taken.copy()
        n, idx, dist = ctx.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t1)
        rn, ridx, rdist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t2)
        assert n == rn and np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and np.array_equal(t1, t2)
        assert n > 0


@pytest.mark.gpu
def test_gpu_search_candidates_edge_cases(slamgpu, oracle):
    """Integer keypoint coordinates (many equal y: the Y order is libstdc++'s std::sort on both sides), 20 000 keypoints,
    radii from 0 (strict '<': nothing inside) to most of the image, queries outside the image."""
    rng = np.random.default_rng(21)
    nk, nq = 20000, 600
    kx = rng.integers(0, 640, nk).astype(np.float32); ky = rng.integers(0, 480, nk).astype(np.float32)
    koct = rng.integers(0, 8, nk).astype(np.int32)
    kdesc = rng.integers(0, 2 ** 32, (nk, 8), dtype=np.uint32)
    src = rng.integers(0, nk, nq)
    qx = (kx[src] + rng.integers(-2, 3, nq)).astype(np.float32); qy = (ky[src] + rng.integers(-2, 3, nq)).astype(np.float32)
    qx[:20] = -500.0; qy[20:40] = 5000.0                                 # far outside
    qr = rng.choice(np.array([0.0, 0.5, 1.0, 3.0, 25.0, 400.0], np.float32), nq)
    qdesc = kdesc[src].copy()
    for i in range(nq):
        for b in rng.integers(0, 256, int(rng.integers(0, 30))):
            qdesc[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    qlevel = rng.integers(0, 8, nq).astype(np.int32)
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        assert np.array_equal(slamgpu.feature_index(kx, ky), oracle.feature_index(kx, ky))
        for kw in (dict(mode=0, thr=50), dict(mode=0, thr=100, qlevel=qlevel)):
            n, idx, dist = ctx.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, **kw)
            rn, ridx, rdist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, **kw)
            assert n == rn and np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and n > 50, kw
        t1, t2 = np.zeros(nk, np.uint8), np.zeros(nk, np.uint8)
        n, idx, dist = ctx.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t1)
        rn, ridx, rdist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t2)
        assert n == rn and np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and np.array_equal(t1, t2)
        assert (idx[qr == 0.0] == -1).all() and (idx[:40] == -1).all()


@pytest.mark.gpu
def test_gpu_search_consumption_forces_rescans(slamgpu, oracle):
    """Many queries compete for few keypoints inside one radius: the truncated top-4 lists run dry and the exact
    rescan path decides."""
    rng = np.random.default_rng(5)
    nk, nq = 40, 200
    kx = rng.uniform(100, 110, nk).astype(np.float32)
    ky = rng.uniform(100, 110, nk).astype(np.float32)
    koct = (np.arange(nk) % 8).astype(np.int32)
    base = rng.integers(0, 2 ** 32, 8, dtype=np.uint32)
    kdesc = np.tile(base, (nk, 1))
    for i in range(nk):                                   # distances 0 .. 19 from the common query descriptor
        for b in rng.choice(256, i // 2, replace=False):
            kdesc[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    qdesc = np.tile(base, (nq, 1))
    qx = np.full(nq, 105, np.float32)
    qy = np.full(nq, 105, np.float32)
    qr = np.full(nq, 30, np.float32)
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        t1, t2 = np.zeros(nk, np.uint8), np.zeros(nk, np.uint8)
        n, idx, dist = ctx.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t1)
        rn, ridx, rdist = oracle.search_candidates(kx, ky, koct, kdesc, qx, qy, qr, qdesc, mode=1, thr=100, taken=t2)
        assert n == rn and np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and np.array_equal(t1, t2)
        assert n > 8 and ctx.rescans() > 0
        # empty inputs
        n, idx, _ = ctx.search_candidates(kx[:0], ky[:0], koct[:0], kdesc[:0], qx, qy, qr, qdesc, mode=0, thr=50)
        assert n == 0 and (idx == -1).all()


@pytest.mark.gpu
def test_gpu_medoid_bit_exact(slamgpu, oracle, synth):
    rng = np.random.default_rng(21)
    sizes = [1, 2, 3, 4, 5, 8, 13, 31, 32, 33, 64, 100, 0, 257, 1024] + rng.integers(1, 40, 200).tolist()
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    desc = rng.integers(0, 2 ** 32, (offs[-1], 8), dtype=np.uint32)
    # observations of one map point are noisy copies of one descriptor (the realistic case: many equal medians)
    for s in range(20, len(sizes)):
        base = desc[offs[s]].copy()
        for i in range(offs[s], offs[s + 1]):
            desc[i] = base
            for b in rng.integers(0, 256, int(rng.integers(0, 12))):
                desc[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        got = ctx.medoid(desc, offs)
        assert np.array_equal(got, oracle.medoid(desc, offs))
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.medoid(rng.integers(0, 2 ** 32, (1025, 8), dtype=np.uint32), np.array([0, 1025], np.int64))


def test_oracle_bow_transform_matches_python(oracle, synth):
    v = synth.random_vocabulary(4, 3, 7, ragged=True)
    rng = np.random.default_rng(8)
    desc = np.concatenate([v["node_desc"][rng.integers(0, len(v["node_word"]), 40)], rng.integers(0, 2 ** 32, (20, 8), dtype=np.uint32)])
    for levels_up in (0, 1, 2, 5):
        word, weight, node = oracle.bow_transform(v, desc, levels_up)
        for f in range(len(desc)):
            cur, level, nid = 0, 0, 0
            while v["child_off"][cur + 1] > v["child_off"][cur]:
                level += 1
                ch = v["child_ids"][v["child_off"][cur]:v["child_off"][cur + 1]]
                dists = [_popc(desc[f], v["node_desc"][c]) for c in ch]
                cur = int(ch[int(np.argmin(dists))])          # argmin: first minimum, like DBoW2's strict '<'
                if level == v["levels"] - levels_up:
                    nid = cur
            assert word[f] == v["node_word"][cur] and node[f] == (0 if v["levels"] - levels_up <= 0 else nid)
            assert weight[f] == v["node_weight"][cur]


@pytest.mark.gpu
@pytest.mark.parametrize("branching,levels,ragged,seed", [(10, 4, False, 31), (7, 3, True, 32), (40, 2, False, 33), (2, 9, True, 34)])
def test_gpu_bow_transform_bit_exact(slamgpu, oracle, synth, branching, levels, ragged, seed):
    v = synth.random_vocabulary(branching, levels, seed, ragged=ragged)
    rng = np.random.default_rng(seed)
    n_nodes = len(v["node_word"])
    desc = np.concatenate([v["node_desc"][rng.integers(0, n_nodes, 1500)], rng.integers(0, 2 ** 32, (500, 8), dtype=np.uint32)])
    for i in range(1500):                                   # noisy copies of vocabulary nodes + unrelated descriptors
        for b in rng.integers(0, 256, int(rng.integers(0, 10))):
            desc[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        voc = slamgpu.Vocabulary(ctx, v)
        for levels_up in (0, 1, 4):
            word, weight, node = voc.transform(desc, levels_up)
            rw, rwt, rn = oracle.bow_transform(v, desc, levels_up)
            assert np.array_equal(word, rw) and np.array_equal(weight, rwt) and np.array_equal(node, rn), levels_up
        assert len(np.unique(word)) > 10
        # the nodes feed the node-bucketed matcher (matchForLoopClosures as the reference runs it)
        dA, dB = desc[:1000], desc[500:1500]
        aA = rng.uniform(0, 360, 1000).astype(np.float32); aB = aA[::-1].copy()
        nA, nB = voc.transform(dA, 1)[2], voc.transform(dB, 1)[2]
        n, m = ctx.match_bow(dA, aA, nA, dB, aB, nB, check_orientation=False)
        rn_, rm_ = oracle.match_bow(dA, aA, nA, dB, aB, nB, check_orientation=False)
        assert n == rn_ and np.array_equal(m, rm_) and n > 100
        voc.close()
        bad = dict(v)
        bad["child_ids"] = v["child_ids"].copy()
        bad["child_ids"][0] = 0                              # a child that points back at the root
        with pytest.raises(slamgpu.SlamGpuError):
            slamgpu.Vocabulary(ctx, bad)


# ---- BowVector, BowIndex::add / remove / getBowSimilar (bow_index.cpp:44-176) -------------------------------------
def _py_bow_vector(word, weight):
    v = {}
    for w, x in zip(word.tolist(), weight.tolist()):
        if x > 0:
            v[w] = v[w] + x if w in v else x
    keys = sorted(v)
    norm = 0.0
    for k in keys:
        norm += abs(v[k])
    return np.array(keys, np.uint32), np.array([v[k] / norm if norm > 0 else v[k] for k in keys], np.float64)


def _py_score(qw, qv, dw, dv):
    q = dict(zip(qw.tolist(), qv.tolist()))
    s, common = 0.0, 0
    for w, x in zip(dw.tolist(), dv.tolist()):          # ascending word order
        if w in q:
            s += abs(q[w] - x) - abs(q[w]) - abs(x)
            common += 1
    return common, np.float32(-s / 2.0)


def _py_similar(db, qw, qv, self_key, min_ratio, score_ratio):
    """db: {(map, kf): (words, values)}; python restatement of bow_index.cpp:95-176 (no score ties assumed)."""
    rows = []
    for key in sorted(db):
        if key == self_key:
            continue
        c, sc = _py_score(qw, qv, *db[key])
        if c:
            rows.append((key, c, sc))
    if not rows:
        return []
    min_common = int(np.float32(min_ratio) * np.float32(max(r[1] for r in rows)))
    sim = sorted([(r[0], r[2]) for r in rows if r[1] > min_common], key=lambda t: -t[1])
    if not sim:
        return []
    min_score = np.float32(sim[0][1] * np.float32(score_ratio))
    return [t for t in sim if not t[1] < min_score]


def test_oracle_bow_vector_kat(oracle):
    w, v = oracle.bow_vector(np.array([5, 3, 5, 7, 3], np.int32), np.array([1.0, 2.0, 3.0, 0.0, 0.5]))
    assert w.tolist() == [3, 5] and v.tolist() == [2.5 / 6.5, 4.0 / 6.5]        # weight 0 = stopped word, dropped
    w, v = oracle.bow_vector(np.zeros(0, np.int32), np.zeros(0))
    assert len(w) == 0
    rng = np.random.default_rng(12)
    word = rng.integers(0, 300, 2000).astype(np.int32)
    weight = np.where(rng.random(2000) < 0.1, 0.0, rng.uniform(0.01, 9.0, 2000))
    w, v = oracle.bow_vector(word, weight)
    pw, pv = _py_bow_vector(word, weight)
    assert np.array_equal(w, pw) and np.array_equal(v, pv)
    assert abs(v.sum() - 1.0) < 1e-12


def test_oracle_bow_similar_kat_and_python(oracle, synth):
    idx = oracle.BowIndex(10)
    a = (np.array([1, 4, 7], np.uint32), np.array([0.5, 0.25, 0.25]))
    b = (np.array([1, 4, 8], np.uint32), np.array([0.5, 0.25, 0.25]))
    c = (np.array([2, 9], np.uint32), np.array([0.5, 0.5]))
    idx.add(0, 10, *a); idx.add(0, 11, *b); idx.add(1, 3, *c)
    m, k, s = idx.similar(*a, self_key=(0, 99), min_in_common_ratio=0.0, score_ratio=0.0)
    assert list(zip(m.tolist(), k.tolist())) == [(0, 10), (0, 11)] and s.tolist() == [1.0, 0.75]   # identical -> 1
    m, k, s = idx.similar(*a, self_key=(0, 10), min_in_common_ratio=0.0, score_ratio=0.0)
    assert list(zip(m.tolist(), k.tolist())) == [(0, 11)]                                          # self excluded
    m, k, s = idx.similar(*a, self_key=(0, 99), min_in_common_ratio=0.8, score_ratio=0.0)
    assert k.tolist() == [10]                                   # 3 common > (unsigned)(0.8 * 3) = 2; 2 common is not
    m, k, s = idx.similar(*a, self_key=(0, 99), min_in_common_ratio=0.0, score_ratio=0.8)
    assert k.tolist() == [10]                                   # 0.75 < 0.8 * 1.0
    idx.remove(0, 10)
    m, k, s = idx.similar(*a, self_key=(0, 99), min_in_common_ratio=0.0, score_ratio=0.0)
    assert k.tolist() == [11]
    assert len(idx.similar(np.array([0, 3], np.uint32), np.array([0.5, 0.5]))[0]) == 0           # no shared word
    idx.close()

    vecs = synth.random_bow_vectors(80, 600, 60, 77)
    db = {}
    idx = oracle.BowIndex(600)
    for i, (w, v) in enumerate(vecs):
        key = (i % 3, 100 - i)
        db[key] = (w, v)
        idx.add(key[0], key[1], w, v)
    for key in [(0, 100), (1, 99), (2, 98)]:
        idx.remove(*key); del db[key]
    for qi in (5, 17, 40, 63):
        key = (qi % 3, 100 - qi)
        for mr, sr in ((0.8, 0.75), (0.3, 0.2), (0.0, 0.0)):
            m, k, s = idx.similar(*db[key], self_key=key, min_in_common_ratio=mr, score_ratio=sr)
            want = _py_similar(db, *db[key], key, mr, sr)
            assert s.tolist() == [float(t[1]) for t in want]
            # equal scores: std::sort's order of ties is libstdc++'s, python's sort is stable -> compare per score
            assert sorted(zip(s.tolist(), m.tolist(), k.tolist())) == sorted((float(t[1]), t[0][0], t[0][1]) for t in want)
            assert len(want) > 0
    idx.close()


@pytest.mark.gpu
def test_gpu_bow_vector_bit_exact(slamgpu, oracle, synth):
    rng = np.random.default_rng(41)
    v = synth.random_vocabulary(10, 3, 5)
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        voc = slamgpu.Vocabulary(ctx, v)
        for n, n_words in ((0, 10), (1, 10), (5, 2), (777, 50), (2000, 1000), (4096, 100000), (4096, 3)):
            word = rng.integers(0, n_words, n).astype(np.int32)
            weight = np.where(rng.random(n) < 0.1, 0.0, rng.uniform(1e-3, 9.0, n))
            w, x = voc.bow_vector(word, weight)
            rw, rx = oracle.bow_vector(word, weight)
            assert np.array_equal(w, rw) and np.array_equal(x, rx), (n, n_words)
        w, x = voc.bow_vector(np.arange(50, dtype=np.int32), np.zeros(50))            # every word stopped
        assert len(w) == 0
        # the real chain: descriptors -> transform -> BowVector
        desc = v["node_desc"][rng.integers(0, len(v["node_word"]), 2000)]
        word, weight, _ = voc.transform(desc, 2)
        w, x = voc.bow_vector(word, weight)
        rw, rx = oracle.bow_vector(*oracle.bow_transform(v, desc, 2)[:2])
        assert np.array_equal(w, rw) and np.array_equal(x, rx) and len(w) > 100
        with pytest.raises(slamgpu.SlamGpuError):
            voc.bow_vector(np.zeros(4097, np.int32), np.ones(4097))
        voc.close()


@pytest.mark.gpu
def test_gpu_bow_vector_batch_bit_exact(slamgpu, oracle, synth):
    """Many keyframes per launch (ragged, with empty and all-stopped keyframes): each equals the single-keyframe result."""
    rng = np.random.default_rng(43)
    v = synth.random_vocabulary(10, 3, 5)
    sizes = [0, 1, 2000, 37, 4096, 512, 0, 1999, 300]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    word = rng.integers(0, 700, int(offs[-1])).astype(np.int32)
    weight = np.where(rng.random(len(word)) < 0.1, 0.0, rng.uniform(1e-3, 9.0, len(word)))
    weight[offs[3]:offs[4]] = 0.0                                     # keyframe 3: every word stopped
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        voc = slamgpu.Vocabulary(ctx, v)
        got = voc.bow_vector_batch(word, weight, offs)
        assert len(got) == len(sizes)
        for k in range(len(sizes)):
            rw, rx = oracle.bow_vector(word[offs[k]:offs[k + 1]], weight[offs[k]:offs[k + 1]])
            assert np.array_equal(got[k][0], rw) and np.array_equal(got[k][1], rx), k
        assert len(got[3][0]) == 0 and len(got[2][0]) > 500
        with pytest.raises(slamgpu.SlamGpuError):
            voc.bow_vector_batch(np.zeros(4097, np.int32), np.ones(4097), np.array([0, 4097], np.int64))
        voc.close()


@pytest.mark.gpu
def test_gpu_bow_similar_bit_exact(slamgpu, oracle, synth):
    vocab_size = 5000
    vecs = synth.random_bow_vectors(400, vocab_size, 300, 55)
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        db = slamgpu.BowDatabase(ctx, 420, 400)
        idx = oracle.BowIndex(vocab_size)
        assert len(db.similar(*vecs[0])[0]) == 0                              # empty database
        keys = []
        for i, (w, v) in enumerate(vecs):
            key = (i % 2, 1000 - i)
            keys.append(key)
            db.add(key[0], key[1], w, v); idx.add(key[0], key[1], w, v)
        # twins (equal scores: the std::sort order of ties must agree), removals, slot reuse
        for j, i in enumerate((3, 3, 50, 51)):
            db.add(7, j, *vecs[i]); idx.add(7, j, *vecs[i]); keys.append((7, j))
        for key in (keys[10], keys[11], keys[200]):
            db.remove(*key); idx.remove(*key)
        db.remove(9, 9)                                                       # unknown keyframe: no-op
        for j, i in enumerate((10, 200)):
            db.add(8, j, *vecs[i][:2]); idx.add(8, j, *vecs[i][:2])
        assert len(db) == 403
        n_results = []
        for qi in (0, 3, 50, 123, 399):
            for self_key in (keys[qi], (-1, -1)):
                for mr, sr in ((0.8, 0.75), (0.5, 0.3), (0.0, 0.0)):
                    got = db.similar(*vecs[qi], self_key=self_key, min_in_common_ratio=mr, score_ratio=sr)
                    ref = idx.similar(*vecs[qi], self_key=self_key, min_in_common_ratio=mr, score_ratio=sr)
                    for g, r in zip(got, ref):
                        assert np.array_equal(g, r), (qi, self_key, mr, sr)
                    n_results.append(len(got[0]))
        assert max(n_results) > 100 and min(n_results) >= 1
        full = db.similar(*vecs[3], min_in_common_ratio=0.0, score_ratio=0.0)
        part = db.similar(*vecs[3], min_in_common_ratio=0.0, score_ratio=0.0, capacity=5)
        assert len(part[0]) == 5 and np.array_equal(part[1], full[1][:5])
        lonely = (np.array([vocab_size + 5], np.uint32), np.array([1.0]))
        assert len(db.similar(*lonely)[0]) == 0                               # shares no word with anybody
        with pytest.raises(slamgpu.SlamGpuError):
            db.add(0, 5000, np.array([4, 3], np.uint32), np.array([0.5, 0.5]))   # words must ascend
        with pytest.raises(slamgpu.SlamGpuError):
            db.add(keys[0][0], keys[0][1], *vecs[0])                          # already present
        db.close(); idx.close()

