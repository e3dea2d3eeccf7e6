"""Columnar observation store and the streaming PlanarDetections-JSON converter (host logic, no GPU):
round trip, the reference's min_corners_per_view filter (facades/intrinsics.cpp:45-47), schema extras
that must be skipped, malformed documents."""
import json

import numpy as np
import pytest

from calibration_b200 import capi


def planar_detections(rng, n_images, sensor="cam0", counts=None):
    """A document following schemas/calib_dataset.schema.json (PlanarDetections, pipeline/dataset.h:15-39)."""
    images = []
    for k in range(n_images):
        n = int(counts[k]) if counts is not None else int(rng.integers(4, 30))
        pts = [{"x": float(rng.uniform(0, 1280)), "y": float(rng.uniform(0, 720)), "id": i,
                "local_x": float(rng.uniform(-0.2, 0.2)), "local_y": float(rng.uniform(-0.2, 0.2)), "local_z": 0.0} for i in range(n)]
        images.append({"file": f"img_{k:04d}.png", "points": pts})
    return {"image_directory": '/data/"quoted" dir', "feature_type": "planar", "algo_version": "1.2", "params_hash": "deadbeef",
            "sensor_id": sensor, "tags": ["synthetic", "x,y"], "metadata": {"nested": {"images": [1, 2, {"points": []}], "e": 1e-3}, "flag": True, "none": None},
            "images": images}


def expected_columns(docs, min_corners):
    off, cam, x, y, u, v = [0], [], [], [], [], []
    for c, doc in enumerate(docs):
        for img in doc["images"]:
            if len(img["points"]) < min_corners:
                continue
            for p in img["points"]:
                x.append(p["local_x"]); y.append(p["local_y"]); u.append(p["x"]); v.append(p["y"])
            off.append(len(x)); cam.append(c)
    return np.array(off), np.array(cam), np.array(x), np.array(y), np.array(u), np.array(v)


@pytest.mark.parametrize("min_corners", [0, 12])
def test_planar_json_to_columns(tmp_path, min_corners):
    rng = np.random.default_rng(4)
    docs = [planar_detections(rng, 25, "left"), planar_detections(rng, 31, "right")]
    paths = []
    for i, d in enumerate(docs):
        p = tmp_path / f"cam{i}.json"
        p.write_text(json.dumps(d, indent=i))   # one compact, one pretty-printed
        paths.append(str(p))
    out = str(tmp_path / "obs.calobs")
    nv, no = capi.Dataset.from_planar_json(paths, out, min_corners_per_view=min_corners)
    off, cam, x, y, u, v = expected_columns(docs, min_corners)
    assert nv == len(cam) and no == len(x)
    with capi.Dataset(out) as ds:
        assert (ds.n_views, ds.n_obs, ds.n_cams) == (len(cam), len(x), 2)
        assert np.array_equal(ds.view_offset, off) and np.array_equal(ds.view_cam, cam)
        # json.dumps writes shortest round-trip reprs and strtod reads them back exactly: bit-equal columns
        assert np.array_equal(ds.x, x) and np.array_equal(ds.y, y) and np.array_equal(ds.u, u) and np.array_equal(ds.v, v)
        for a in (ds.x, ds.y, ds.u, ds.v):
            assert a.ctypes.data % 64 == 0


def test_write_open_round_trip_and_empty_views(tmp_path):
    rng = np.random.default_rng(1)
    lens = np.array([5, 0, 88, 1, 54])
    off = np.concatenate([[0], np.cumsum(lens)])
    cols = [rng.normal(size=off[-1]) for _ in range(4)]
    path = str(tmp_path / "a.calobs")
    capi.Dataset.write(path, off, [0, 1, 1, 2, 0], *cols, n_cams=3)
    with capi.Dataset(path) as ds:
        assert ds.n_cams == 3 and np.array_equal(ds.view_offset, off) and ds.view_cam.tolist() == [0, 1, 1, 2, 0]
        for a, b in zip((ds.x, ds.y, ds.u, ds.v), cols):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("text, msg", [
    ('{"sensor_id": "c"}', "'images' is required"),
    ('{"sensor_id": "c", "images": [{"file": "a", "points": [{"x": 1, "y": 2, "local_x": 3}]}]}', "required"),
    ('{"sensor_id": "c", "images": [{"file": "a", "points": [{"x": 1, "y": 2, "local_x": 3, "local_y": "4"}]}]}', "expected a number"),
    ('{"sensor_id": "c", "images": [{"file": "a", "points": [', "expected"),
    ('{"sensor_id": "c", "images": []} trailing', "trailing"),
])
def test_malformed_documents_are_rejected(tmp_path, text, msg):
    p = tmp_path / "bad.json"; p.write_text(text)
    with pytest.raises(ValueError, match=msg):
        capi.Dataset.from_planar_json([str(p)], str(tmp_path / "o.calobs"))


def test_open_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.bin"; p.write_bytes(bytes(256))
    with pytest.raises(RuntimeError, match="not a CALOBS01 file"):
        capi.Dataset(str(p))
    with pytest.raises(RuntimeError, match="cannot open"):
        capi.Dataset(str(tmp_path / "missing.calobs"))


def _header(n_views, n_obs, n_cams, off_view_offset, off_view_cam, off_x, stride_obs):
    import struct
    return b"CALOBS01" + struct.pack("<qqiiqqqq", n_views, n_obs, n_cams, 0, off_view_offset, off_view_cam, off_x, stride_obs)


def test_open_rejects_crafted_and_truncated_headers(tmp_path):
    """cal_dataset_open validates every section against the file size before it dereferences anything: a header whose
    counts or offsets point outside the file, a truncated file, non-monotone CSR offsets and camera ids out of range
    are all reported, never dereferenced."""
    rng = np.random.default_rng(2)
    off = np.array([0, 3, 7])
    cols = [rng.normal(size=7) for _ in range(4)]
    good = tmp_path / "g.calobs"
    capi.Dataset.write(str(good), off, [0, 1], *cols, n_cams=2)
    blob = bytearray(good.read_bytes())
    with capi.Dataset(str(good)) as ds:
        assert ds.n_views == 2

    def opens(data, name):
        p = tmp_path / name; p.write_bytes(bytes(data)); return capi.Dataset(str(p))

    crafted = {
        "huge_n_views": _header(2 ** 40, 0, 1, 64, 64, 64, 0) + bytes(64),                     # the 128-byte reproducer
        "n_views_past_eof": _header(1000, 7, 2, 64, 128, 192, 64) + bytes(blob[64:]),
        "stride_overflow": _header(2, 7, 2, 64, 128, 192, 2 ** 62) + bytes(blob[64:]),
        "off_x_past_eof": _header(2, 7, 2, 64, 128, 2 ** 40, 64) + bytes(blob[64:]),
        "cam_overlaps_offsets": _header(2, 7, 2, 64, 72, 192, 64) + bytes(blob[64:]),
        "negative_obs": _header(2, -7, 2, 64, 128, 192, 64) + bytes(blob[64:]),
        "truncated": bytes(blob[:len(blob) - 40]),
    }
    for name, data in crafted.items():
        with pytest.raises(RuntimeError, match="not a CALOBS01 file"):
            opens(data, name)
    import struct
    bad = bytearray(blob); struct.pack_into("<q", bad, 64 + 8, 9)                               # view_offset = [0, 9, 7]
    with pytest.raises(RuntimeError, match="monotonically"):
        opens(bad, "non_monotone")
    hdr = struct.unpack_from("<qqiiqqqq", blob, 8)
    bad = bytearray(blob); struct.pack_into("<i", bad, hdr[5], 5)                               # view_cam[0] = 5 >= n_cams
    with pytest.raises(RuntimeError, match="out of range"):
        opens(bad, "cam_range")


@pytest.mark.parametrize("text, msg", [
    ('{"images": [{"points": [{"x": nan, "y": 2, "local_x": 3, "local_y": 4}]}]}', "expected a number"),
    ('{"images": [{"points": [{"x": inf, "y": 2, "local_x": 3, "local_y": 4}]}]}', "expected a number"),
    ('{"images": [{"points": [{"x": 0x10, "y": 2, "local_x": 3, "local_y": 4}]}]}', "malformed number"),
    ('{"images": [{"points": [{"x": +1, "y": 2, "local_x": 3, "local_y": 4}]}]}', "expected a number"),
    ('{"images": [{"points": [{"x": 1e999, "y": 2, "local_x": 3, "local_y": 4}]}]}', "out of range"),
    ('{"images": [{"points": [{"x": 1., "y": 2, "local_x": 3, "local_y": 4}]}]}', "malformed number"),
    ('{"meta": ' + "[" * 100000 + "]" * 100000 + ', "images": []}', "nesting too deep"),
])
def test_numbers_follow_the_json_grammar_and_nesting_is_bounded(tmp_path, text, msg):
    p = tmp_path / "bad.json"; p.write_text(text)
    with pytest.raises(ValueError, match=msg):
        capi.Dataset.from_planar_json([str(p)], str(tmp_path / "o.calobs"))
