"""the reference's eye-tracking database (video_database.mat, MATLAB v7.3 = HDF5) through the package's own minimal
HDF5 reader, and the fixation fixtures for all 33 clips that were extracted with it (tests/golden/make_golden_fixations.py)."""
import os

import numpy as np
import pytest

import _oracle as o
from compressd_domain_saliencyprediction_b200 import mat73, metrics

GOLD = os.path.join(o.ROOT, "tests", "golden")
DB = "/root/reference/video_database.mat"


def test_fixture_covers_all_33_clips_of_the_database():
    db = metrics.load_fixation_database(os.path.join(GOLD, "fixations_all.npz"))
    assert len(db["names"]) == 33 and db["names"][0] == "Kimono" and db["names"][20] == "BasketballPass" and db["names"][-1] == "PeopleOnStreet"
    assert db["fixdata"].shape == (26914, 6)
    vid = db["fixdata"][:, 1]
    assert np.isnan(vid).sum() < 50                                        # a few all-NaN rows separate the subjects' records
    assert set(np.unique(vid[np.isfinite(vid)]).astype(int)) == set(range(1, 34))
    for k, name in enumerate(db["names"]):
        w, h = db["size"][k]
        assert (w, h) in [(1920, 1080), (832, 480), (416, 240), (1280, 720), (1024, 768), (1280, 800)], (name, w, h)
        assert 100 <= db["frames"][k] <= 700 and 20 <= db["fps"][k] <= 60, (name, db["frames"][k], db["fps"][k])
        f = db["fixdata"][db["fixdata"][:, 1] == k + 1]
        inside = np.mean((f[:, 4] > 0) & (f[:, 4] <= w) & (f[:, 5] > 0) & (f[:, 5] <= h))
        assert len(f) > 300 and inside > 0.9, (name, len(f), inside)


@pytest.mark.parametrize("name", ["BasketballPass", "BasketballDrill", "BasketballDrive"])
def test_per_picture_assignment_reproduces_the_clip_fixtures(name):
    """main.m:331-344 on the 33-clip table gives exactly the per-picture fixtures the GPU AUC / NSS tests use."""
    db = metrics.load_fixation_database(os.path.join(GOLD, "fixations_all.npz"))
    fx = np.load(os.path.join(GOLD, f"fixations_{name}.npz"))
    xy, off = metrics.fixations_per_picture(db, name, fx["pocs"])
    assert np.array_equal(xy, fx["xy"]) and np.array_equal(off, fx["offsets"])
    xy_all, off_all = metrics.fixations_per_picture(db, name)
    assert len(off_all) == db["frames"][db["names"].index(name)] + 1 and off_all[-1] == len(xy_all) > len(xy)


def test_hdf5_reader_on_the_reference_database():
    if not os.path.exists(DB):
        pytest.skip("needs /root/reference/video_database.mat (build container)")
    raw = mat73.load(DB)["video_database"]
    assert raw["fixdata_fields"][:4] == ["SubjectIndex", "VideoIndex", "Timestamp", "GazeDuration"]
    assert raw["videos_info"]["name_list"] == raw["video_names_list"] and len(raw["subj_names_list"]) >= 20
    db = metrics.load_fixation_database(DB)
    fx = metrics.load_fixation_database(os.path.join(GOLD, "fixations_all.npz"))
    assert db["names"] == fx["names"]
    for k in ("size", "frames", "fps", "fixdata"):
        assert np.array_equal(np.asarray(db[k], np.float32), np.asarray(fx[k], np.float32), equal_nan=True), k


def test_hdf5_reader_rejects_other_files(tmp_path):
    p = tmp_path / "x.mat"
    p.write_bytes(b"MATLAB 5.0 MAT-file" + b"\x00" * 2000)
    with pytest.raises(mat73.Mat73Error):
        mat73.load(str(p))
