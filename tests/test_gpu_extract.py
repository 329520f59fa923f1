"""GPU parity tests (through the C ABI) of the extraction path against the CPU oracle:
pyramid pixels, FAST candidates, keypoints after the quadtree distribution (set AND order),
orientation angles and rBRIEF descriptors -- all bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_params(po, w, h, **kw):
    return po.make_params(w, h, **kw)


@pytest.mark.parametrize("w,h,seed", [(640, 480, 1000), (1280, 720, 4000), (333, 247, 5), (97, 131, 6)])
def test_pyramid_bit_exact(slamgpu, oracle, synth, w, h, seed):
    levels = 8 if min(w, h) > 200 else 4
    img = synth.frame(w, h, seed)
    with slamgpu.Context(w, h, levels=levels, max_frames=1) as ctx:
        ctx.pyramid_update(img)
        lv, bl = oracle.pyramid(_oracle_params(oracle, w, h, levels=levels), img)
        for l in range(levels):
            got = ctx.get_level(0, l)
            assert got.shape == lv[l].shape
            assert np.array_equal(got, lv[l]), "pyramid level %d differs in %d px" % (l, (got != lv[l]).sum())
            gb = ctx.get_blurred_level(0, l)
            assert np.array_equal(gb, bl[l]), "blurred level %d differs in %d px" % (l, (gb != bl[l]).sum())


def test_pyramid_golden_crc(slamgpu, synth, golden):
    import zlib
    for name, (w, h, seed) in {"vga": (640, 480, 1000), "hd": (1280, 720, 4000)}.items():
        img = synth.frame(w, h, seed)
        with slamgpu.Context(w, h, max_frames=1) as ctx:
            ctx.pyramid_update(img)
            crc = golden["cv2"]["pyr_%s_crc" % name]
            for l in range(8):
                assert zlib.crc32(ctx.get_level(0, l).tobytes()) == crc[0, l], (name, l)
                assert zlib.crc32(ctx.get_blurred_level(0, l).tobytes()) == crc[1, l], (name, l)


@pytest.mark.parametrize("factor,levels", [(2.0, 3), (1.5, 4), (1.1, 6)])
def test_pyramid_other_scale_factors(slamgpu, oracle, synth, factor, levels):
    # factor 2.0 makes cv::resize switch to INTER_AREA silently
    img = synth.frame(320, 240, 31)
    with slamgpu.Context(320, 240, levels=levels, scale_factor=factor, max_frames=1) as ctx:
        ctx.pyramid_update(img)
        lv, bl = oracle.pyramid(oracle.make_params(320, 240, levels=levels, scale_factor=factor), img)
        for l in range(levels):
            assert np.array_equal(ctx.get_level(0, l), lv[l]), l
            assert np.array_equal(ctx.get_blurred_level(0, l), bl[l]), l


def test_pyramid_degenerate_and_batch(slamgpu, oracle, synth):
    deg = synth.degenerate_frames(640, 480)
    imgs = np.stack(list(deg.values()) + [synth.frame(640, 480, 77)])
    with slamgpu.Context(640, 480, max_frames=len(imgs)) as ctx:
        ctx.pyramid_update(imgs)
        for f in range(len(imgs)):
            lv, bl = oracle.pyramid(oracle.make_params(640, 480), imgs[f])
            for l in range(8):
                assert np.array_equal(ctx.get_level(f, l), lv[l]), (f, l)
                assert np.array_equal(ctx.get_blurred_level(f, l), bl[l]), (f, l)


def test_pyramid_strided_host_input(slamgpu, oracle, synth):
    big = np.zeros((480, 700), np.uint8)
    img = synth.frame(640, 480, 12)
    big[:, :640] = img
    with slamgpu.Context(640, 480, max_frames=1) as ctx:
        ctx.pyramid_update(big[:, :640])
        lv, _ = oracle.pyramid(oracle.make_params(640, 480), img)
        assert np.array_equal(ctx.get_level(0, 3), lv[3])


@pytest.mark.parametrize("w,h,seed,maxkp", [(640, 480, 1000, 2000), (640, 480, 1001, 1000), (1280, 720, 4000, 2000),
                                            (333, 247, 5, 500), (640, 480, 1002, 60)])
def test_detection_bit_exact(slamgpu, oracle, synth, w, h, seed, maxkp):
    img = synth.frame(w, h, seed)
    p = oracle.make_params(w, h, max_keypoints=maxkp)
    _, _, _, budgets = oracle.geometry(p)
    lv, _ = oracle.pyramid(p, img)
    with slamgpu.Context(w, h, max_keypoints=maxkp, max_frames=1) as ctx:
        assert np.array_equal(ctx.budgets, budgets)
        ctx.pyramid_update(img)
        ctx.detect()
        for l in range(8):
            (ox, oy, orr), (cx, cy, cr) = oracle.detect_level(lv[l], int(budgets[l]), with_candidates=True)
            gx, gy, gr = ctx.detected(0, l, candidates=True)
            got = sorted(zip(gx.tolist(), gy.tolist(), gr.tolist()))
            ref = sorted(zip(cx.tolist(), cy.tolist(), cr.tolist()))
            assert got == ref, "level %d: FAST candidates differ (%d vs %d)" % (l, len(got), len(ref))
            kx, ky, kr = ctx.detected(0, l)
            assert kx.tolist() == ox.tolist() and ky.tolist() == oy.tolist() and kr.tolist() == orr.tolist(), \
                "level %d: keypoints after distribution differ (%d vs %d)" % (l, len(kx), len(ox))


def test_detection_degenerate_frames(slamgpu, oracle, synth):
    deg = synth.degenerate_frames(640, 480)
    rng = np.random.default_rng(3)
    deg["noise"] = rng.integers(0, 256, (480, 640), dtype=np.uint8)   # very many corners
    imgs = np.stack(list(deg.values()))
    p = oracle.make_params(640, 480)
    _, _, _, budgets = oracle.geometry(p)
    with slamgpu.Context(640, 480, max_frames=len(imgs)) as ctx:
        ctx.pyramid_update(imgs)
        ctx.detect()
        for f in range(len(imgs)):
            lv, _ = oracle.pyramid(p, imgs[f])
            for l in range(8):
                ox, oy, orr = oracle.detect_level(lv[l], int(budgets[l]))
                kx, ky, kr = ctx.detected(f, l)
                assert kx.tolist() == ox.tolist() and ky.tolist() == oy.tolist() and kr.tolist() == orr.tolist(), (f, l)


def _assert_same_extraction(got, ref):
    assert got["n"] == ref["n"], (got["n"], ref["n"])
    assert np.array_equal(got["level_counts"], ref["level_counts"])
    assert np.array_equal(got["lvl_x"], ref["lvl_x"]) and np.array_equal(got["lvl_y"], ref["lvl_y"])
    assert np.array_equal(got["octave"], ref["octave"])
    assert np.array_equal(got["track_id"], ref["track_id"])
    assert np.array_equal(got["x"], ref["x"]) and np.array_equal(got["y"], ref["y"])
    # north star: angles within 1e-4 rad; the scalar fastAtan2 restatement is in fact bit-equal
    dang = np.abs(got["angle"].astype(np.float64) - ref["angle"].astype(np.float64))
    assert (np.deg2rad(dang) <= 1e-4).all()
    assert np.array_equal(got["angle"], ref["angle"]), "angles not bit-equal: max diff %g deg" % dang.max()
    flips = int((got["desc"] != ref["desc"]).any(axis=1).sum())
    assert flips == 0, "%d descriptors differ" % flips


@pytest.mark.parametrize("w,h,seed,maxkp", [(640, 480, 1000, 1000), (640, 480, 2000, 2000), (1280, 720, 4000, 2000)])
def test_extract_bit_exact(slamgpu, oracle, synth, w, h, seed, maxkp):
    img = synth.frame(w, h, seed)
    with slamgpu.Context(w, h, max_keypoints=maxkp, max_frames=1) as ctx:
        got = ctx.detect_and_extract(img)[0]
        ref = oracle.extract(oracle.make_params(w, h, max_keypoints=maxkp), img)
        _assert_same_extraction(got, ref)


@pytest.mark.parametrize("w,h,levels,factor,maxkp", [(752, 480, 8, 1.2, 1500), (641, 479, 5, 1.2, 800), (128, 96, 3, 1.2, 200),
                                                     (320, 240, 1, 1.2, 300), (800, 600, 6, 1.3, 5000), (1920, 1080, 8, 1.2, 4000),
                                                     (96, 64, 8, 1.2, 100), (64, 64, 4, 1.2, 50), (3840, 2160, 8, 1.2, 8000),
                                                     (1920, 480, 8, 1.2, 1500), (480, 1920, 8, 1.2, 1500), (4096, 64, 8, 1.2, 500),
                                                     (64, 4096, 8, 1.2, 500), (640, 480, 8, 1.2, 1), (640, 480, 8, 1.2, 8), (640, 480, 8, 1.2, 30)])
def test_extract_unusual_geometries(slamgpu, oracle, synth, w, h, levels, factor, maxkp):
    """Image sizes that are not multiples of the tile / cell sizes, a single-level pyramid, another scale factor,
    budgets above and below the usual 2000, a 4K frame, 4:1 and 64:1 aspect ratios (more initial quadtree nodes than
    the level's budget)."""
    img = synth.frame(w, h, 8000 + w)
    with slamgpu.Context(w, h, levels=levels, scale_factor=factor, max_keypoints=maxkp, max_frames=1) as ctx:
        got = ctx.detect_and_extract(img)[0]
        ref = oracle.extract(oracle.make_params(w, h, levels=levels, scale_factor=factor, max_keypoints=maxkp), img)
        _assert_same_extraction(got, ref)
        assert got["n"] > 0


@pytest.mark.parametrize("ini,mn", [(40, 10), (15, 15), (7, 1), (120, 30), (254, 200)])
def test_extract_other_fast_thresholds(slamgpu, oracle, synth, ini, mn):
    """FAST thresholds other than the upstream 20 / 7: equal thresholds (no fallback pass), very low ones (almost every
    pixel scores: the scored-pixel list overflows into the map scan), very high ones (few or no corners)."""
    img = synth.frame(640, 480, 3100 + ini)
    with slamgpu.Context(640, 480, max_keypoints=1500, ini_fast_thr=ini, min_fast_thr=mn, max_frames=1) as ctx:
        got = ctx.detect_and_extract(img)[0]
    ref = oracle.extract(oracle.make_params(640, 480, max_keypoints=1500, ini_fast_thr=ini, min_fast_thr=mn), img)
    _assert_same_extraction(got, ref)
    assert (got["n"] > 0) == (ini < 254)


def test_extract_batch_matches_single_frames(slamgpu, oracle, synth):
    imgs = synth.frames(640, 480, 6, 2100)
    with slamgpu.Context(640, 480, max_frames=6) as ctx:
        got = ctx.detect_and_extract(imgs)
        for f in range(6):
            _assert_same_extraction(got[f], oracle.extract(oracle.make_params(640, 480), imgs[f]))
        # a partial batch on the same context
        got2 = ctx.detect_and_extract(imgs[3:5])
        for f in range(2):
            _assert_same_extraction(got2[f], got[3 + f])


def test_extract_pipeline_chunks_agree(slamgpu, oracle, synth):
    """sg_extract pipelines H2D / kernels / D2H over chunks of the batch; any chunking gives the same result."""
    imgs = synth.frames(640, 480, 7, 2300)
    ref = [oracle.extract(oracle.make_params(640, 480), imgs[f]) for f in range(7)]
    with slamgpu.Context(640, 480, max_frames=7) as ctx:
        for chunk in (1, 2, 3, 7, 64):
            ctx.set_pipeline_chunk(chunk)
            got = ctx.detect_and_extract(imgs)
            for f in range(7):
                _assert_same_extraction(got[f], ref[f])
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.set_pipeline_chunk(0)


def test_extract_streaming_batches_in_flight(slamgpu, oracle, synth):
    """sg_extract_submit / sg_extract_wait: batches queued back to back on disjoint frame slots of one context
    (the copies of one batch run under the kernels of the other) give the same keypoints as the oracle."""
    B, n_batches = 5, 6
    imgs = synth.frames(640, 480, B * n_batches, 2600)
    p = oracle.make_params(640, 480)
    ref = {f: oracle.extract(p, imgs[f]) for f in (0, 4, 7, 13, 16, 22, 29)}
    with slamgpu.Context(640, 480, max_frames=2 * B + 1) as ctx:
        ctx.set_pipeline_chunk(2)
        pins = [slamgpu.PinnedArray((B, 480, 640), np.uint8) for _ in range(2)]
        outs = [ctx.alloc_outputs(B, pinned=True) for _ in range(2)]
        tickets = [None, None]
        got = {}

        def collect(b):
            ctx.extract_wait(tickets[b % 2])
            for f, d in enumerate(ctx._split(outs[b % 2][0], B)):
                got[b * B + f] = d

        for b in range(n_batches):
            if b >= 2:
                collect(b - 2)
            pins[b % 2].array[...] = imgs[b * B:(b + 1) * B]
            # odd batches sit at an odd slot offset: the frame-slot base is arbitrary
            tickets[b % 2] = ctx.extract_submit(pins[b % 2].array, (b % 2) * (B + 1), outs[b % 2][1])
        collect(n_batches - 2)
        collect(n_batches - 1)
        for f, r in ref.items():
            _assert_same_extraction(got[f], r)
        # the synchronous call still works on the same context afterwards
        _assert_same_extraction(ctx.detect_and_extract(imgs[7])[0], ref[7])
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.extract_submit(pins[0].array, 2 * B, outs[0][1])        # slots [10, 15) exceed the 11 of the context
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.extract_wait(7)                                          # never issued
        t = ctx.extract_submit(pins[0].array, 0, outs[0][1])
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.extract_submit(pins[1].array, 3, outs[1][1])            # slots [3, 8) overlap the batch in flight
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.detect_and_extract(imgs[0])                              # the synchronous call uses slot 0 too
        ctx.extract_wait(t)
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.extract_wait(t)                                          # already waited for
        _assert_same_extraction(ctx._split(outs[0][0], B)[2], ref[22])      # pins[0] still holds batch 4 = frames 20..24


def test_extract_with_tracker_features(slamgpu, oracle, synth):
    img = synth.frame(640, 480, 3100)
    rng = np.random.default_rng(5)
    tracks = np.stack([rng.uniform(-5, 645, 80), rng.uniform(-5, 485, 80)], axis=1).astype(np.float32)
    tracks[:4] = [[19.4, 19.4], [19.6, 18.4], [620.5, 460.5], [621.5, 461.5]]   # margin / half-even rounding cases
    ids = (np.arange(80) * 3 + 11).astype(np.int32)
    for level in (0, 2):
        with slamgpu.Context(640, 480, max_frames=1, max_tracks=128, track_level=level) as ctx:
            got = ctx.detect_and_extract(img, tracks=[tracks], track_ids=[ids])[0]
            ref = oracle.extract(oracle.make_params(640, 480), img, tracks=tracks, track_ids=ids, track_level=level)
            assert (ref["track_id"] >= 0).sum() > 20
            _assert_same_extraction(got, ref)


def test_extract_degenerate_frames(slamgpu, oracle, synth):
    deg = synth.degenerate_frames(640, 480)
    imgs = np.stack(list(deg.values()))
    with slamgpu.Context(640, 480, max_frames=len(imgs)) as ctx:
        got = ctx.detect_and_extract(imgs)
        for f in range(len(imgs)):
            _assert_same_extraction(got[f], oracle.extract(oracle.make_params(640, 480), imgs[f]))
    assert got[0]["n"] == 0 and got[1]["n"] == 0   # empty result is legal (orb_extractor.cpp:137)


def test_extract_device_resident_full_batch(slamgpu, oracle, synth):
    """BASELINE config 2 shape (256 frames) through the device-resident entry point; every frame of
    the batch is independent, so frame f of the batch must equal the oracle on frame f (checked on a
    sample) and repeated frames must give identical results (checked on all)."""
    base = synth.frames(640, 480, 8, 5000)
    idx = np.arange(256) % 8
    imgs = base[idx]
    with slamgpu.Context(640, 480, max_frames=256) as ctx:
        buf = ctx.device_buffer(imgs.nbytes).upload(imgs)
        ctx.extract_device(buf.ptr, 640, 640 * 480, 256)
        got = ctx.extract_download(256)
        buf.free()
    for f in range(8):
        _assert_same_extraction(got[f], oracle.extract(oracle.make_params(640, 480), base[f]))
    for f in range(8, 256):
        g, r = got[f], got[f % 8]
        assert g["n"] == r["n"] and np.array_equal(g["desc"], r["desc"]) and np.array_equal(g["angle"], r["angle"])
        assert np.array_equal(g["x"], r["x"]) and np.array_equal(g["y"], r["y"])


def test_extract_256_distinct_frames_all_against_oracle(slamgpu, oracle):
    """The timed batch of bench.py (BASELINE configs[1]: 256 DISTINCT 640x480 frames, 2000 keypoints) through the three entry
    points -- device resident, host buffers, streaming -- every frame compared with the oracle, every field bit for bit."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    import bench
    imgs = bench.make_frames(256, 10000)
    assert len({imgs[f].tobytes() for f in range(256)}) == 256
    p = oracle.make_params(640, 480, max_keypoints=2000)
    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        refs = list(ex.map(lambda f: oracle.extract(p, imgs[f]), range(256)))
    with slamgpu.Context(640, 480, max_keypoints=2000, max_frames=256) as ctx:
        buf = ctx.device_buffer(imgs.nbytes).upload(imgs)
        ctx.extract_device(buf.ptr, 640, 640 * 480, 256)
        dev = ctx.extract_download(256)
        host = ctx.detect_and_extract(imgs)
        arrs, ks = ctx.alloc_outputs(256, pinned=True)
        ctx.extract_wait(ctx.extract_submit(imgs, 0, ks))
        buf.free()
        cap = ctx.cap
        for f in range(256):
            _assert_same_extraction(dev[f], refs[f])
            _assert_same_extraction(host[f], refs[f])
            n = refs[f]["n"]
            assert int(arrs["count"][f]) == n
            for k in ("x", "y", "angle", "octave", "desc"):
                assert np.array_equal(arrs[k][f, :n], refs[f][k]), (f, k)
            assert cap >= n
    assert sum(r["n"] for r in refs) > 256 * 1900


def test_extract_rejects_bad_frame_counts_with_tracks(slamgpu):
    """n_frames is validated before anything is sized by it (tracker-point staging included)."""
    with slamgpu.Context(640, 480, max_frames=2, max_tracks=8) as ctx:
        imgs = np.zeros((3, 480, 640), np.uint8)
        tracks = [np.array([[100.0, 100.0]], np.float32)] * 3
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.detect_and_extract(imgs, tracks=tracks)
        ok = ctx.detect_and_extract(imgs[:2], tracks=tracks[:2])      # the context still works afterwards
        assert len(ok) == 2


def test_errors(slamgpu):
    with pytest.raises(slamgpu.SlamGpuError):
        slamgpu.Context(32, 32)
    with slamgpu.Context(640, 480, max_frames=2) as ctx:
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.detect()   # before any pyramid
        with pytest.raises(slamgpu.SlamGpuError):
            ctx.pyramid_update(np.zeros((3, 480, 640), np.uint8))   # more frames than the context holds
