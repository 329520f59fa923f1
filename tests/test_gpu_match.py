"""GPU parity tests (through the C ABI) of the Hamming / brute-force matching path against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(slamgpu):
    c = slamgpu.Context(640, 480, max_frames=1)
    yield c
    c.close()


def test_hamming_golden(ctx, golden):
    r = golden["ref"]
    assert np.array_equal(ctx.hamming(r["hamm_a"], r["hamm_b"]), r["hamm_d"])
    a = np.array([[0xffffffff, 0, 1, 2, 3, 4, 5, 6]], np.uint32)
    assert ctx.hamming(a, np.zeros((1, 8), np.uint32))[0] == 41   # SURVEY 8c known answer


@pytest.mark.parametrize("n,seed,kw", [
    (2000, 1, {}), (2000, 2, dict(flip_bits=60, keep=0.9)), (500, 3, dict(flip_bits=8, keep=1.0, angle_jitter=40.0)),
    (1237, 4, dict(flip_bits=30, keep=0.5)), (33, 5, {}), (1, 6, {}),
])
def test_match_bruteforce_bit_exact(ctx, oracle, synth, n, seed, kw):
    dA, aA, dB, aB = synth.correlated_descriptors(n, seed, **kw)
    for check in (True, False):
        got_n, got = ctx.match_bruteforce(dA, aA, dB, aB, check_orientation=check)
        ref_n, ref = oracle.match_bruteforce(dA, aA, dB, aB, check_orientation=check)
        assert got_n == ref_n and np.array_equal(got, ref), (n, seed, check, got_n, ref_n)
    if n >= 500:
        assert ref_n > 0


def test_match_duplicates_force_rescans(ctx, oracle):
    """Many identical B descriptors: the top-4 lists of later rows are all consumed, so the exact
    full-row rescan path decides; ties must resolve to the lowest index like the strict '<' scan."""
    rng = np.random.default_rng(11)
    base = rng.integers(0, 2 ** 32, (20, 8), dtype=np.uint32)
    dA = np.repeat(base, 40, axis=0)
    dB = np.repeat(base, 40, axis=0)[rng.permutation(800)]
    flips = rng.integers(0, 3, len(dB))
    for i, k in enumerate(flips):
        for b in rng.integers(0, 256, k):
            dB[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    aA = rng.uniform(0, 360, len(dA)).astype(np.float32)
    aB = rng.uniform(0, 360, len(dB)).astype(np.float32)
    for ratio in (0.8, 1.0, 0.5):
        got_n, got = ctx.match_bruteforce(dA, aA, dB, aB, ratio=ratio, check_orientation=False)
        ref_n, ref = oracle.match_bruteforce(dA, aA, dB, aB, ratio=ratio, check_orientation=False)
        assert got_n == ref_n and np.array_equal(got, ref), ratio
    assert ctx.rescans() > 0


def test_match_duplicates_force_rescans_in_large_launches(ctx, slamgpu, oracle):
    """The same clustered sets through a launch with many pairs: that path keeps 4-entry lists (few pairs get 16), so
    its exact rescans must be exercised separately."""
    rng = np.random.default_rng(12)
    base = rng.integers(0, 2 ** 32, (30, 8), dtype=np.uint32)
    sets_d, sets_a = [], []
    for s in range(3):
        d = np.repeat(base, 20, axis=0)[rng.permutation(600)]
        for i, k in enumerate(rng.integers(0, 3, len(d))):
            for b in rng.integers(0, 256, k):
                d[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
        sets_d.append(d)
        sets_a.append(rng.uniform(0, 360, len(d)).astype(np.float32))
    db = slamgpu.DescriptorDB(ctx, np.stack(sets_d), np.stack(sets_a))
    pairs = np.array([(i % 3, (i // 3) % 3) for i in range(180)], np.int32)       # 180 pairs x 3 row tiles: the 4-entry path
    n, m = db.match_pairs(pairs, check_orientation=False)
    assert ctx.rescans() > 0
    ref = {(i, j): oracle.match_bruteforce(sets_d[i], sets_a[i], sets_d[j], sets_a[j], check_orientation=False)
           for i in range(3) for j in range(3)}
    for k, (i, j) in enumerate(pairs.tolist()):
        assert n[k] == ref[(i, j)][0] and np.array_equal(m[k], ref[(i, j)][1]), k
    db.close()


def test_match_ragged_and_empty(ctx, oracle, synth):
    dA, aA, dB, aB = synth.correlated_descriptors(700, 21)
    for na, nb in [(700, 300), (300, 700), (700, 1), (1, 700), (257, 1025), (0, 10), (10, 0)]:
        got_n, got = ctx.match_bruteforce(dA[:na], aA[:na], dB[:nb], aB[:nb])
        ref_n, ref = oracle.match_bruteforce(dA[:na], aA[:na], dB[:nb], aB[:nb]) if na and nb else (0, np.full(na, -1, np.int32))
        assert got_n == ref_n and np.array_equal(got, ref), (na, nb)


@pytest.mark.parametrize("ratio,thr,dbl", [(0.8, 50, False), (0.8, 50, True), (0.6, 100, False), (1.0, 30, False), (0.95, 256, False)])
def test_match_parameters(ctx, oracle, synth, ratio, thr, dbl):
    dA, aA, dB, aB = synth.correlated_descriptors(900, 31, flip_bits=70, keep=0.8)
    got_n, got = ctx.match_bruteforce(dA, aA, dB, aB, ratio=ratio, thr=thr, ratio_is_double=dbl)
    ref_n, ref = oracle.match_bruteforce(dA, aA, dB, aB, ratio=ratio, thr=thr, ratio_is_double=dbl)
    assert got_n == ref_n and np.array_equal(got, ref)


def test_match_pairs_batched(ctx, slamgpu, oracle, synth):
    """A small database of ragged sets, all ordered pairs in one call."""
    sizes = [500, 1, 777, 0, 2000, 64]
    rng = np.random.default_rng(41)
    root, ang = synth.random_descriptors(1, 2000, 77)
    root, ang = root[0], ang[0]
    desc, angles, offs = [], [], [0]
    for s in sizes:
        sel = rng.permutation(2000)[:s]
        d = root[sel].copy()
        for i in range(s):
            for b in rng.integers(0, 256, rng.integers(0, 25)):
                d[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
        desc.append(d)
        angles.append((ang[sel] + rng.normal(0, 3, s)).astype(np.float32) % np.float32(360))
        offs.append(offs[-1] + s)
    D = np.concatenate(desc)
    A = np.concatenate(angles).astype(np.float32)
    db = slamgpu.DescriptorDB(ctx, D, A, np.array(offs, np.int64))
    pairs = np.array([(i, j) for i in range(len(sizes)) for j in range(len(sizes))], np.int32)
    n, m = db.match_pairs(pairs)
    db.close()
    for k, (i, j) in enumerate(pairs):
        if sizes[i] == 0 or sizes[j] == 0:
            assert n[k] == 0 and (m[k] == -1).all()
            continue
        ref_n, ref = oracle.match_bruteforce(desc[i], angles[i], desc[j], angles[j])
        assert n[k] == ref_n, (i, j)
        assert np.array_equal(m[k, :sizes[i]], ref), (i, j)
        assert (m[k, sizes[i]:] == -1).all()


def test_match_pairs_many_slabs(ctx, slamgpu, oracle, synth):
    """More pairs than one 256-pair slab: the host-buffer call double-buffers the match rows (copy-out of slab i under
    the kernels of slab i+1); rows, counts and the -1 padding must come out as from single calls."""
    dA, aA, dB, aB = synth.correlated_descriptors(300, 91)
    dC, aC, dD, aD = synth.correlated_descriptors(300, 92)
    sets_d, sets_a = [dA, dB, dC[:190], dD], [aA, aB, aC[:190], aD]
    offs = np.concatenate([[0], np.cumsum([len(d) for d in sets_d])]).astype(np.int64)
    db = slamgpu.DescriptorDB(ctx, np.concatenate(sets_d), np.concatenate(sets_a).astype(np.float32), offs)
    rng = np.random.default_rng(5)
    pairs = rng.integers(0, 4, (700, 2)).astype(np.int32)
    n, m = db.match_pairs(pairs)
    n2, _ = db.match_pairs(pairs, want_matches=False)
    assert np.array_equal(n, n2)
    cache = {}
    for k, (i, j) in enumerate(pairs.tolist()):
        if (i, j) not in cache:
            cache[(i, j)] = oracle.match_bruteforce(sets_d[i], sets_a[i], sets_d[j], sets_a[j])
        rn, rm = cache[(i, j)]
        assert n[k] == rn and np.array_equal(m[k, :len(rm)], rm) and (m[k, len(rm):] == -1).all(), k
    assert int(n.max()) > 100
    db.close()


@pytest.mark.parametrize("view", [False, True])
def test_match_from_device_views(slamgpu, oracle, synth, view):
    """The one-GPU extract -> match flow of the stereo config: descriptors never leave the device.  The database is a
    copy (sg_db_create_device) or a view (sg_db_wrap_device) of the extractor's output arrays; set 2f = keypoints of
    frame f, set 2f+1 = the unused tail of its slot."""
    a, b = synth.frame(640, 480, 1200), synth.frame(640, 480, 1201)
    imgs = np.stack([a, synth.shifted_rotated(a), b, np.roll(b, -5, axis=1)])
    with slamgpu.Context(640, 480, max_frames=4) as c:
        buf = c.device_buffer(imgs.nbytes).upload(imgs)
        c.extract_device(buf.ptr, 640, 640 * 480, 4)
        kps = c.extract_download(4)
        counts = np.array([k["n"] for k in kps])
        offs = np.zeros(9, np.int64)
        offs[1::2] = np.arange(4) * c.cap + counts
        offs[2::2] = (np.arange(4) + 1) * c.cap
        v = c.device_views()
        db = slamgpu.DescriptorDB(c, None, None, offsets=offs, device_ptrs=(v.desc, v.angle), view=view)
        pairs = np.array([(0, 2), (4, 6), (6, 4), (2, 2)], np.int32)
        n, m = db.match_pairs(pairs)
        db.close()
        buf.free()
    for k, (i, j) in enumerate(pairs // 2):
        rn, rm = oracle.match_bruteforce(kps[i]["desc"], kps[i]["angle"], kps[j]["desc"], kps[j]["angle"])
        assert n[k] == rn and np.array_equal(m[k, :len(rm)], rm), (k, view)
    assert n[0] > 20 and n[3] == counts[1]


@pytest.mark.parametrize("nA,nB", [(8000, 6000), (65535, 300), (300, 65535)])
def test_match_large_sets(ctx, slamgpu, oracle, nA, nB):
    """Sets far beyond the usual 2000 features, up to the 65535 the 16-bit index fields allow; with nA > nB many A rows
    compete for the same B feature (uniqueness pressure on the sequential walk)."""
    rng = np.random.default_rng(nA + nB)
    dB = rng.integers(0, 2 ** 32, (nB, 8), dtype=np.uint32)
    aB = rng.uniform(0, 360, nB).astype(np.float32)
    src = rng.integers(0, nB, nA)
    dA = dB[src].copy()
    bits = rng.integers(0, 256, (nA, 40))
    keep = rng.integers(0, 40, nA)
    for i in range(nA):
        for b in bits[i, :keep[i]]:
            dA[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    aA = ((aB[src] + rng.normal(0, 4, nA)) % 360).astype(np.float32)
    n, m = ctx.match_bruteforce(dA, aA, dB, aB)
    rn, rm = oracle.match_bruteforce(dA, aA, dB, aB)
    assert n == rn and np.array_equal(m, rm) and n >= min(nA, nB) // 2
    if nB == 65535:
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.match_bruteforce(dA, aA, np.concatenate([dB, dB[:1]]), np.concatenate([aB, aB[:1]]))   # 65536 features


def test_match_extracted_frames(slamgpu, oracle, synth):
    """End to end on real descriptors: frame vs shifted/rotated frame (BASELINE config 3 inputs (i))."""
    img = synth.frame(640, 480, 1000)
    img2 = synth.shifted_rotated(img)
    with slamgpu.Context(640, 480, max_frames=2) as c:
        k1, k2 = c.detect_and_extract(np.stack([img, img2]))
        got_n, got = slamgpu.match_for_loop_closures(c, k1, k2)
    ref_n, ref = oracle.match_bruteforce(k1["desc"], k1["angle"], k2["desc"], k2["angle"])
    assert got_n == ref_n and np.array_equal(got, ref)
    assert ref_n > 20


def test_match_properties_full_size(ctx, slamgpu, synth):
    """Size-independent properties at BASELINE config 3 size (2000 x 2000), many pairs per call."""
    d, a = synth.random_descriptors(8, 2000, 99)
    db = slamgpu.DescriptorDB(ctx, d, a)
    # a set against itself: every feature's best is itself at distance 0 -> identity, all in bin 0
    pairs = np.array([(i, i) for i in range(8)], np.int32)
    n, m = db.match_pairs(pairs)
    assert (n == 2000).all() and (m == np.arange(2000)).all()
    # unrelated random sets: distances concentrate near 128, nothing survives thr 50
    pairs = np.array([(i, j) for i in range(8) for j in range(8) if i != j], np.int32)
    n, m = db.match_pairs(pairs)
    assert (n == 0).all() and (m == -1).all()
    # counts-only call agrees
    n2, _ = db.match_pairs(pairs, want_matches=False)
    assert np.array_equal(n, n2)
    db.close()


@pytest.mark.parametrize("seed,n_nodes", [(41, 1), (42, 12), (43, 97)])
def test_match_bow_node_buckets(ctx, oracle, synth, seed, n_nodes):
    """matchForLoopClosures as the reference runs it: DBoW2 node buckets (keyframe_matcher.cpp:65-146), map-point
    eligibility filters, one angle histogram over all nodes."""
    rng = np.random.default_rng(seed)
    dA, aA, dB, aB = synth.correlated_descriptors(900, seed)
    # a vocabulary stand-in: node = a few descriptor bits (noisy copies mostly land in the same node), some features in no node
    bits = lambda d: ((d[:, 0] ^ d[:, 3]) % np.uint32(n_nodes)).astype(np.int32) * 3 + 5
    nodeA, nodeB = bits(dA), bits(dB)
    nodeA[rng.random(900) < 0.05] = -1
    nodeB[rng.random(900) < 0.05] = -1
    eligA = (rng.random(900) < 0.9).astype(np.uint8)
    eligB = (rng.random(900) < 0.9).astype(np.uint8)
    for eA, eB in ((None, None), (eligA, eligB)):
        n, m = ctx.match_bow(dA, aA, nodeA, dB, aB, nodeB, eA, eB)
        rn, rm = oracle.match_bow(dA, aA, nodeA, dB, aB, nodeB, eA, eB)
        assert n == rn and np.array_equal(m, rm), (n, rn)
    if n_nodes == 1:
        keep = nodeA >= 0
        assert n > 50
    # disjoint node sets and empty inputs
    n, m = ctx.match_bow(dA, aA, np.zeros(900, np.int32), dB, aB, np.ones(900, np.int32))
    assert n == 0 and (m == -1).all()
    n, m = ctx.match_bow(dA[:0], aA[:0], nodeA[:0], dB, aB, nodeB)
    assert n == 0 and len(m) == 0


def _two_view_scene(seed, n=700, n_nodes=9):
    """Two keyframes looking at random 3-D points: bearings, the essential matrix the reference builds
    (create_E_21(kf2.R, kf2.t, kf1.R, kf1.t), openvslam/essential_solver.cc:157-162), noisy descriptor copies."""
    rng = np.random.default_rng(seed)
    def rot(ax, ang):
        ax = ax / np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    R1, t1 = np.eye(3), np.zeros(3)
    R2, t2 = rot(rng.normal(size=3), 0.1), np.array([0.5, 0.05, -0.02])
    X = rng.uniform(-3, 3, (n, 3)) + np.array([0, 0, 8.0])
    b1 = (R1 @ X.T).T + t1; b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
    b2 = (R2 @ X.T).T + t2; b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    # E = create_E_21(rot_1w = R2, trans_1w = t2, rot_2w = R1, trans_2w = t1)
    R21 = R1 @ R2.T
    t21 = -R21 @ t2 + t1
    E = np.array([[0, -t21[2], t21[1]], [t21[2], 0, -t21[0]], [-t21[1], t21[0], 0]]) @ R21
    dA = rng.integers(0, 2 ** 32, (n, 8), dtype=np.uint32)
    dB = dA.copy()
    for i in range(n):
        for b in rng.integers(0, 256, int(rng.integers(0, 40))):
            dB[i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    # duplicates: several kf2 features with the same descriptor (ties: the last one wins), some off the epipolar plane
    dB[1::9] = dB[0::9][:len(dB[1::9])]
    wrong = rng.random(n) < 0.25
    b2[wrong] = rng.normal(size=(int(wrong.sum()), 3)); b2[wrong] /= np.linalg.norm(b2[wrong], axis=1, keepdims=True)
    perm = rng.permutation(n)
    dB, b2 = dB[perm], b2[perm]
    aA = rng.uniform(0, 360, n).astype(np.float32)
    aB = ((aA + rng.normal(0, 4, n)) % 360).astype(np.float32)[perm]
    octA = rng.integers(0, 8, n).astype(np.int32)
    node = lambda d: ((d[:, 1] >> 7) % np.uint32(n_nodes)).astype(np.int32)
    # nodes from bits the noise rarely touches would be ideal; a vocabulary stand-in: share the node of the source feature mostly
    nodeA = rng.integers(0, n_nodes, n).astype(np.int32)
    nodeB = nodeA.copy()
    nodeB[rng.random(n) < 0.1] = rng.integers(0, n_nodes)
    nodeB = nodeB[perm]
    return dA, aA, octA, b1, nodeA, dB, aB, b2, nodeB, E


@pytest.mark.parametrize("seed,n,n_nodes", [(61, 700, 9), (62, 1500, 1), (63, 300, 40)])
def test_match_triangulation_bit_exact(ctx, oracle, seed, n, n_nodes):
    """matchForTriangulationDBoW (keyframe_matcher.cpp:160-293): node buckets, last-wins ties, epipolar test, uniqueness."""
    dA, aA, octA, b1, nodeA, dB, aB, b2, nodeB, E = _two_view_scene(seed, n, n_nodes)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    rng = np.random.default_rng(seed)
    eligA = (rng.random(n) < 0.8).astype(np.uint8)
    eligB = (rng.random(n) < 0.8).astype(np.uint8)
    for thr_deg in (0.2, 2.0):
        for eA, eB in ((None, None), (eligA, eligB)):
            got = ctx.match_triangulation(dA, aA, octA, b1, nodeA, dB, aB, b2, nodeB, E, sf, eA, eB, residual_deg_thr=thr_deg)
            ref = oracle.match_triangulation(dA, aA, octA, b1, nodeA, dB, aB, b2, nodeB, E, sf, eA, eB, residual_deg_thr=thr_deg)
            assert got[0] == ref[0] and np.array_equal(got[1], ref[1]), (thr_deg, got[0], ref[0])
    assert ref[0] > n // 20


def test_match_triangulation_exhausted_lists_are_rescanned(ctx, oracle):
    """Many identical kf2 descriptors that fail the epipolar test in front of the one that passes: the top-4 list runs dry
    and the exact row scan decides."""
    rng = np.random.default_rng(9)
    n = 64
    base = rng.integers(0, 2 ** 32, 8, dtype=np.uint32)
    dA = np.tile(base, (n, 1)); dB = np.tile(base, (n, 1))
    b1 = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    E = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])      # pure x translation: epipolar plane normal = (0, -b2z, b2y)
    b2 = np.tile(np.array([0.0, 0.6, 0.8]), (n, 1))            # off the plane of b1
    ok = rng.permutation(n)[:20]
    b2[ok] = [0.3, 0.0, 0.954]                                  # on it
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    aA = np.zeros(n, np.float32); aB = np.zeros(n, np.float32)
    octA = np.zeros(n, np.int32); node = np.zeros(n, np.int32)
    sf = np.ones(8, np.float32)
    got = ctx.match_triangulation(dA, aA, octA, b1, node, dB, aB, b2, node, E, sf)
    ref = oracle.match_triangulation(dA, aA, octA, b1, node, dB, aB, b2, node, E, sf)
    assert got[0] == ref[0] == 20 and np.array_equal(got[1], ref[1])
    assert ctx.rescans() > 0


def test_database_may_outlive_its_context(slamgpu):
    """C callers may destroy the context before a database created on it (only the Python wrapper orders this):
    sg_destroy releases the database's device memory and detaches it, sg_db_destroy afterwards only frees the handle."""
    import ctypes as C
    L = slamgpu.lib()
    ctx = slamgpu.Context(640, 480, max_frames=1)
    d, a = np.zeros((2, 64, 8), np.uint32), np.zeros((2, 64), np.float32)
    offs = np.array([0, 64, 128], np.int64)
    db = C.c_void_p()
    assert L.sg_db_create(ctx._h, d.ctypes.data, a.ctypes.data, offs.ctypes.data, 2, C.byref(db)) == 0
    h, ctx._h = ctx._h, None          # destroy the context behind the wrapper's back, database still alive
    L.sg_destroy(h)
    L.sg_db_destroy(db)               # must not touch the dead context
    with slamgpu.Context(640, 480, max_frames=1) as c2:      # the device is still healthy
        n, m = c2.match_bruteforce(d[0], a[0], d[1], a[1])
        assert n >= 0
