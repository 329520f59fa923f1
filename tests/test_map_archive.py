"""Map archives: the cereal binary layout of MapDB::serialize (mapdb.hpp:83-98) restated INDEPENDENTLY
here with struct.pack -- cereal's published binary-archive rules: raw arithmetic values, uint64 element counts, uint32
shared-pointer ids with the top bit on first occurrence, KeyPoint::serialize writing `octave` twice (key_point.hpp:22-25) --
and checked against the C++ reader / writer of slam-module_b200/host/map_archive.cpp: the reader must recover every field
and the writer must reproduce the file byte for byte.  cereal and the parent project's Eigen / cv::Vec3b adapters are
absent: PARITY UNPINNED (no archive written by the reference exists offline)."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "slam-module_b200" / "host"


def _exe():
    subprocess.check_call(["make", "-C", str(HOST), "all"], stdout=subprocess.DEVNULL)
    return HOST / "archive_main"


def _mat(m, fixed_header):
    m = np.asarray(m, np.float64)
    hdr = struct.pack("<ii", *m.shape) if fixed_header else b""
    return hdr + m.T.astype("<f8").tobytes()            # column major


def _vec3(v, fmt, fixed_header):
    return (struct.pack("<ii", 3, 1) if fixed_header else b"") + struct.pack("<3" + fmt, *v)


def build_archive(fixed_header=False, seed=0, n_kf=3, n_kp=40):
    """-> (bytes, description dict) of a small map: keyframes with keypoints, map points observed by two keyframes."""
    rng = np.random.default_rng(seed)
    next_ptr = [1]
    out = []
    info = {"kfs": [], "mps": []}
    descs = rng.integers(0, 2 ** 32, (n_kf, n_kp, 8), dtype=np.uint32)
    descs[1] = descs[0]                                   # keyframe 1 re-observes keyframe 0's features
    flip = rng.integers(0, 256, (n_kp, 6))
    for i in range(n_kp):
        for b in flip[i]:
            descs[1, i, b >> 5] ^= np.uint32(1 << (int(b) & 31))
    angles = rng.uniform(0, 360, (n_kf, n_kp)).astype(np.float32)
    angles[1] = (angles[0] + 10) % 360

    def keypoint(k, i):
        x, y = float(np.float32(rng.uniform(0, 640))), float(np.float32(rng.uniform(0, 480)))
        octave = int(rng.integers(0, 8))
        bearing = rng.normal(size=3)
        bearing /= np.linalg.norm(bearing)
        if i == 0:
            info["kfs"][-1]["kp0"] = (x, y, float(angles[k, i]), octave, float(bearing[2]), int(descs[k, i, 7]))
        return (struct.pack("<fffii", x, y, float(angles[k, i]), octave, octave) + _vec3(bearing, "d", fixed_header)
                + descs[k, i].astype("<u4").tobytes())

    def shared_block(k):
        cam = ("pinhole %d" % k).encode()
        b = struct.pack("<Q", len(cam)) + cam
        b += struct.pack("<Q", n_kp) + b"".join(keypoint(k, i) for i in range(n_kp))
        colors = rng.integers(0, 256, (n_kp, 3), dtype=np.uint8)
        b += struct.pack("<Q", n_kp) + colors.tobytes()
        if k == 1:                                        # a stereo point cloud on one keyframe, null on the others
            pid = next_ptr[0]; next_ptr[0] += 1
            cloud = rng.normal(size=(5, 3)).astype(np.float32)
            b += struct.pack("<I", pid | 0x80000000) + struct.pack("<Q", 5) + b"".join(_vec3(p, "f", fixed_header) for p in cloud)
        else:
            b += struct.pack("<I", 0)
        return b

    body = struct.pack("<Q", n_kf)
    for k in range(n_kf):
        kf_id = 10 + 5 * k
        info["kfs"].append({"id": kf_id, "prev": kf_id - 5 if k else -1, "next": kf_id + 5 if k + 1 < n_kf else -1, "t": 0.5 * k,
                            "tx": 0.25 * k})
        body += struct.pack("<i", kf_id)                                      # map key
        pid = next_ptr[0]; next_ptr[0] += 1
        body += struct.pack("<I", pid | 0x80000000)                           # shared_ptr<Keyframe>, first occurrence
        spid = next_ptr[0]; next_ptr[0] += 1
        body += struct.pack("<I", spid | 0x80000000) + shared_block(k)        # shared_ptr<KeyframeShared>
        body += struct.pack("<iii", kf_id, info["kfs"][-1]["prev"], info["kfs"][-1]["next"])
        body += struct.pack("<Q", 2) + struct.pack("<iiii", 0, 100 + k, 3, 200 + k)          # keyPointToTrackId
        mps = [i if (k < 2 and i % 2 == 0) else -1 for i in range(n_kp)]
        body += struct.pack("<Q", n_kp) + struct.pack("<%di" % n_kp, *mps)                   # mapPoints
        body += struct.pack("<Q", n_kp) + rng.uniform(0.5, 9, n_kp).astype("<f4").tobytes()  # keyPointDepth
        pose = np.eye(4); pose[0, 3] = 0.25 * k
        body += _mat(pose, fixed_header) + _mat(np.eye(4), fixed_header) + _mat(rng.normal(size=(3, 6)), fixed_header)
        body += struct.pack("<d", 0.5 * k) + struct.pack("<B", 1)
    mp_ids = [i for i in range(n_kp) if i % 2 == 0]
    body += struct.pack("<Q", len(mp_ids))
    for i in mp_ids:
        pos = rng.normal(size=3)
        status = int(rng.integers(0, 4))
        info["mps"].append({"id": i, "status": status, "pos": pos.tolist()})
        body += struct.pack("<i", i)                                          # map key
        body += struct.pack("<iii", i, -1, status) + _vec3(pos, "d", fixed_header) + _vec3(rng.normal(size=3), "f", fixed_header)
        body += struct.pack("<ff", 0.5, 30.0) + descs[0, i].astype("<u4").tobytes()
        body += struct.pack("<Q", 2) + struct.pack("<iiii", 10, i, 15, i)     # observations: both keyframes, same keypoint index
        body += struct.pack("<i", 10) + bytes([1, 2, 3])
    body += struct.pack("<Q", 1) + struct.pack("<ii", 100, 0)                 # trackIdToMapPoint
    body += struct.pack("<Q", 1) + struct.pack("<ii", 10, 15) + _mat(np.eye(4), fixed_header)   # loopClosureEdges
    body += _mat(np.eye(4), fixed_header) + _mat(np.eye(4), fixed_header)
    body += struct.pack("<qq", 3, 6) + np.zeros((3, 6)).T.astype("<f8").tobytes()             # MatrixXd discardedUncertainty
    body += struct.pack("<d", 12.5) + struct.pack("<iii", n_kp, 10 + 5 * (n_kf - 1), 10 + 5 * (n_kf - 1))
    info["descs"], info["angles"] = descs, angles
    return body, info


@pytest.mark.parametrize("fixed_header", [False, True])
def test_archive_round_trip_is_byte_exact(tmp_path, fixed_header):
    exe = _exe()
    data, info = build_archive(fixed_header, seed=3)
    src, dst = tmp_path / "map.bin", tmp_path / "map_out.bin"
    src.write_bytes(data)
    args = [str(exe), "roundtrip", str(src), str(dst)] + (["fixed"] if fixed_header else [])
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert dst.read_bytes() == data
    lines = r.stdout.splitlines()
    assert lines[0].startswith("keyframes 3 keypoints 120 mappoints 20 tracks 1 edges 1 nextMp 40 lastKf 20 t0 12.500")
    for k, kf in enumerate(info["kfs"]):
        x, y, a, o, bz, d7 = kf["kp0"]
        want = "kf %d prev %d next %d kps 40 camera 'pinhole %d' t %.3f pose03 %.3f full 1 kp0 %.2f %.2f %.2f %d %.3f %08x" % (
            kf["id"], kf["prev"], kf["next"], k, kf["t"], kf["tx"], x, y, a, o, bz, d7)
        assert lines[1 + k] == want
    for j, mp in enumerate(info["mps"]):
        assert lines[4 + j] == "mp %d status %d obs 2 pos %.3f %.3f %.3f ref 10" % (mp["id"], mp["status"], *mp["pos"])


def test_archive_reader_rejects_damage(tmp_path):
    exe = _exe()
    data, _ = build_archive(False, seed=4)
    for name, blob in (("short", data[:len(data) // 2]), ("long", data + b"\0"), ("count", struct.pack("<Q", 1 << 60) + data[8:])):
        f = tmp_path / (name + ".bin")
        f.write_bytes(blob)
        r = subprocess.run([str(exe), "roundtrip", str(f), str(tmp_path / "o.bin")], capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout.startswith("error:"), (name, r.stdout)


@pytest.mark.gpu
def test_archive_rebuilds_descriptor_database(tmp_path, oracle):
    """loadMapDB's rebuild hook (mapper_helpers.cpp:958-993) for the GPU: the loaded keyframes become an sg_db and are matched
    with sg_match_pairs; the result equals the oracle's brute-force matcher on the descriptors the test put in the file."""
    exe = _exe()
    data, info = build_archive(False, seed=5, n_kp=300)
    src = tmp_path / "map.bin"
    src.write_bytes(data)
    r = subprocess.run([str(exe), "match", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("matched ")][0]
    head, tail = line.split(":")
    got = np.array(tail.split(), np.int32)
    n, want = oracle.match_bruteforce(info["descs"][0], info["angles"][0], info["descs"][1], info["angles"][1])
    assert head == "matched %d of 300 between kf 10 and kf 15" % n and np.array_equal(got, want) and n > 100
    assert "feature search rebuilt: 300" in r.stdout
