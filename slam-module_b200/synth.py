"""Deterministic synthetic inputs for tests and benches (numpy only; no dataset, no network).

Frames follow SURVEY.md 8(d): uniform 8-bit noise smoothed with a sigma-1.5 Gaussian, then
filled axis-aligned rectangles of random size 8..40 px and random gray level painted on top
(200 for 640x480, scaled with area).  This yields ~1e4 raw FAST corners over 8 levels, so every
level exceeds its keypoint budget and the quadtree stage is exercised.
"""
import numpy as np


def _smooth(img, sigma=1.5):
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    f = img.astype(np.float32)
    f = np.pad(f, ((0, 0), (r, r)), mode="reflect")
    f = sum(k[i] * f[:, i:i + img.shape[1]] for i in range(2 * r + 1))
    f = np.pad(f, ((r, r), (0, 0)), mode="reflect")
    f = sum(k[i] * f[i:i + img.shape[0], :] for i in range(2 * r + 1))
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def frame(width, height, seed, n_rect=None):
    rng = np.random.default_rng(seed)
    img = _smooth(rng.integers(0, 256, (height, width), dtype=np.uint8))
    if n_rect is None:
        n_rect = int(round(200 * width * height / (640 * 480)))
    for _ in range(n_rect):
        w = int(rng.integers(8, 41))
        h = int(rng.integers(8, 41))
        x = int(rng.integers(0, width - w))
        y = int(rng.integers(0, height - h))
        img[y:y + h, x:x + w] = rng.integers(0, 256)
    return img


def frames(width, height, n, seed0):
    return np.stack([frame(width, height, seed0 + i) for i in range(n)])


def degenerate_frames(width, height):
    """Border / empty-cell cases: all-zero, all-255, vertical step, checkerboard of 16-px squares."""
    z = np.zeros((height, width), np.uint8)
    o = np.full((height, width), 255, np.uint8)
    s = z.copy()
    s[:, width // 2:] = 200
    yy, xx = np.mgrid[0:height, 0:width]
    c = (((yy // 16) + (xx // 16)) % 2 * 180 + 30).astype(np.uint8)
    return {"zeros": z, "ones": o, "step": s, "checker": c}


def shifted_rotated(img, dx=3, dy=2, deg=5.0):
    """Second view of a frame (nearest-neighbour warp) so that ratio/angle tests pass for a
    realistic fraction of features."""
    h, w = img.shape
    a = np.deg2rad(deg)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cx, cy = w / 2.0, h / 2.0
    xs = np.cos(a) * (xx - cx) + np.sin(a) * (yy - cy) + cx - dx
    ys = -np.sin(a) * (xx - cx) + np.cos(a) * (yy - cy) + cy - dy
    xi = np.clip(np.rint(xs).astype(np.int64), 0, w - 1)
    yi = np.clip(np.rint(ys).astype(np.int64), 0, h - 1)
    return np.ascontiguousarray(img[yi, xi])


def random_descriptors(n_sets, n_per_set, seed):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 2 ** 32, (n_sets, n_per_set, 8), dtype=np.uint32)
    a = rng.uniform(0, 360, (n_sets, n_per_set)).astype(np.float32)
    return d, a


def correlated_descriptors(n_per_set, seed, flip_bits=20, keep=0.6, angle_jitter=4.0, rot=30.0):
    """Two descriptor sets where a fraction `keep` of B are noisy copies of A (flip_bits random bit
    flips, common rotation + jitter), shuffled; exercises ratio / uniqueness / angle semantics."""
    rng = np.random.default_rng(seed)
    dA = rng.integers(0, 2 ** 32, (n_per_set, 8), dtype=np.uint32)
    aA = rng.uniform(0, 360, n_per_set).astype(np.float32)
    dB = rng.integers(0, 2 ** 32, (n_per_set, 8), dtype=np.uint32)
    aB = rng.uniform(0, 360, n_per_set).astype(np.float32)
    src = rng.permutation(n_per_set)
    for j in range(n_per_set):
        if rng.random() < keep:
            i = src[j]
            d = dA[i].copy()
            nflip = int(rng.integers(0, flip_bits + 1))
            for b in rng.integers(0, 256, nflip):
                d[b >> 5] ^= np.uint32(1 << (int(b) & 31))
            dB[j] = d
            aB[j] = np.float32((aA[i] - rot + rng.normal(0, angle_jitter)) % 360.0)
    return dA, aA, dB, aB


def random_vocabulary(branching, levels, seed, ragged=False):
    """A DBoW2-shaped vocabulary tree for tests and benches (there is no vocabulary file offline): node 0 is the
    root, every inner node has `branching` children (2..branching when ragged) whose descriptors are noisy copies of
    the parent's, leaves at depth `levels` carry a word id and a weight.  Children are stored breadth first, so the
    ids of a node's children are contiguous, but the arrays use DBoW2's general form (a child-id list per node)."""
    rng = np.random.default_rng(seed)
    desc = [rng.integers(0, 2 ** 32, 8, dtype=np.uint32)]
    depth = [0]
    child_off, child_ids = [0], []
    i = 0
    while i < len(desc):
        if depth[i] < levels:
            k = int(rng.integers(2, branching + 1)) if ragged else branching
            for _ in range(k):
                d = desc[i].copy()
                for b in rng.integers(0, 256, max(2, 48 >> depth[i])):
                    d[b >> 5] ^= np.uint32(1 << (int(b) & 31))
                if rng.random() < 0.05 and child_ids and len(child_ids) > child_off[-1]:
                    d = desc[child_ids[-1]].copy()          # a twin of the previous sibling: distance ties
                child_ids.append(len(desc))
                desc.append(d)
                depth.append(depth[i] + 1)
        child_off.append(len(child_ids))
        i += 1
    n = len(desc)
    child_off = np.asarray(child_off, np.int32)
    is_leaf = np.diff(child_off) == 0
    word = np.full(n, -1, np.int32)
    word[is_leaf] = np.arange(int(is_leaf.sum()), dtype=np.int32)
    weight = np.where(is_leaf, rng.uniform(0.1, 5.0, n), 0.0).astype(np.float64)
    weight[is_leaf & (rng.random(n) < 0.03)] = 0.0         # "stopped" words: DBoW2 skips features with weight 0
    return dict(child_off=child_off, child_ids=np.asarray(child_ids, np.int32), node_desc=np.stack(desc).astype(np.uint32),
                node_weight=weight, node_word=word, levels=levels)


def random_bow_vectors(n_keyframes, vocabulary_size, words_per_keyframe, seed, n_topics=12):
    """Sparse L1-normalised BowVectors for database tests / benches: every keyframe draws most of its words from one
    of `n_topics` overlapping word pools (so that keyframes of a topic share many words) plus uniform noise words.
    -> list of (words ascending uint32, values float64)."""
    rng = np.random.default_rng(seed)
    pools = [rng.choice(vocabulary_size, min(vocabulary_size, 3 * words_per_keyframe), replace=False) for _ in range(n_topics)]
    out = []
    for _ in range(n_keyframes):
        pool = pools[int(rng.integers(0, n_topics))]
        k = int(rng.integers(max(1, words_per_keyframe // 2), words_per_keyframe + 1))
        w = np.unique(np.concatenate([rng.choice(pool, min(len(pool), k), replace=False),
                                      rng.integers(0, vocabulary_size, max(1, k // 8))])).astype(np.uint32)
        v = rng.uniform(0.1, 5.0, len(w)) * rng.integers(1, 4, len(w))
        out.append((w, (v / np.abs(v).sum()).astype(np.float64)))
    return out
