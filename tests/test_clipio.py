"""Row N4: the clip-level loaders / writers (skiing_analysis_pytorch_b200/clipio.py) against golden G10 - fixture files in
the reference's on-disk schemas (tests/golden/io/) together with what the REFERENCE's own loaders returned for them
(triangulation/load.py, fuse/load/load_raw.py) and the files the reference's own writers produced (fuse/save.py,
triangulation/save.py).  Loader outputs must be equal; writer outputs byte-identical."""
import filecmp
from pathlib import Path

import numpy as np
import pytest
import torch

from skiing_analysis_pytorch_b200 import clipio

IO = Path(__file__).resolve().parent / "golden" / "io"


@pytest.mark.parametrize("side", ["left", "right"])
def test_pt_keypoint_loader_matches_reference(golden, side):
    g = golden("g10_io.npz")
    xy, sc = clipio.load_keypoints_pt(IO / f"{side}.pt", source="YOLO")      # normalised coordinates -> pixels via img_shape
    np.testing.assert_array_equal(xy, g[f"{side}_yolo_xy"])
    np.testing.assert_array_equal(sc, g[f"{side}_yolo_score"])
    assert xy.max() > 100.0
    xy, sc = clipio.load_keypoints_pt(IO / f"{side}.pt", source="detectron2")  # pixels; score falls back to keypoints[..., 2]
    np.testing.assert_array_equal(xy, g[f"{side}_d2_xy"])
    np.testing.assert_array_equal(sc, g[f"{side}_d2_score"])
    with pytest.raises(KeyError):
        clipio.load_keypoints_pt(IO / f"{side}.pt", source="openpose")


def test_pt_loader_edge_cases():
    k = torch.rand(4, 17, 2)
    xy, sc = clipio.load_keypoints_pt({"YOLO": {"keypoints": k}}, source="YOLO")      # no frame size: returned as is; no scores anywhere: ones
    np.testing.assert_array_equal(xy, k.numpy())
    assert sc.shape == (4, 17) and (sc == 1).all()
    xy, _ = clipio.load_keypoints_pt({"YOLO": {"keypoints": k}, "img_shape": (100, 200)}, source="YOLO", assume_normalized=False)
    np.testing.assert_array_equal(xy, k.numpy())
    xy, _ = clipio.load_keypoints_pt({"YOLO": {"keypoints": k}, "frames": torch.zeros(4, 100, 200, 3, dtype=torch.uint8)}, source="YOLO")
    np.testing.assert_allclose(xy[..., 0], k.numpy()[..., 0] * 200)
    with pytest.raises(ValueError):
        clipio.load_keypoints_pt({"YOLO": {"keypoints": torch.rand(4, 17)}}, source="YOLO")
    with pytest.raises(ValueError):
        clipio.load_keypoints_pt({"YOLO": {"keypoints": k, "keypoints_score": torch.rand(4, 16)}}, source="YOLO")


def test_sam3d_loaders_match_reference_load_raw(golden):
    g = golden("g10_io.npz")
    Xl, Xr, Ul, Ur = clipio.load_sam3d_pair({"sam_l": IO / "osmo_2_sam_3d_body_outputs.npz", "sam_r": IO / "right"})
    assert Xl.shape == (6, 70, 3) and Ur.shape == (6, 70, 2) and Xl.dtype == np.float64   # truncated to the shorter view
    for got, key in ((Xl, "sam_Xl"), (Xr, "sam_Xr"), (Ul, "sam_Ul"), (Ur, "sam_Ur")):
        np.testing.assert_array_equal(got, g[key])
    p2, p3 = clipio.load_sam3d_sequence(IO / "right")
    assert p3.shape == (7, 70, 3)
    with pytest.raises(FileNotFoundError):
        clipio.load_sam3d_sequence(IO / "nowhere")


def test_writers_are_byte_identical_to_the_references(golden, tmp_path):
    g = golden("g10_io.npz")
    p = clipio.save_sequence_npy(g["seq_to_save"], tmp_path / "o" / "person_smoothed.npy")
    assert filecmp.cmp(p, IO / "ref_out" / "person_smoothed.npy", shallow=False)
    p2 = clipio.save_sequence_npy(torch.from_numpy(g["seq_to_save"]), tmp_path / "o2" / "x.npy")
    assert filecmp.cmp(p2, IO / "ref_out" / "person_smoothed.npy", shallow=False)
    with pytest.raises(ValueError):
        clipio.save_sequence_npy(np.zeros((0, 70, 3)), tmp_path / "o" / "e.npy")
    vp = {"left": "/data/left.mp4", "right": Path("/data/right.mp4")}
    for fmt in ("npy", "csv", "json"):
        files = clipio.save_3d_joints_clip(g["joints_X"], tmp_path / "j", g["joints_R"], g["joints_T"], vp, fmt=fmt, first_frame=40)
        assert len(files) == 3
        for f in files:
            assert filecmp.cmp(f, IO / "ref_out" / "joints" / Path(f).name, shallow=False), f
    with pytest.raises(ValueError):
        clipio.save_3d_joints_clip(g["joints_X"], tmp_path / "j", g["joints_R"], g["joints_T"], vp, fmt="xml")


def test_camera_npz_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    N, C = 5, 2
    K = [rng.normal(size=(C, 3, 3)) for _ in range(N)]
    R = [rng.normal(size=(C, 3, 3)) for _ in range(N)]
    t = [rng.normal(size=(C, 3)) for _ in range(N)]
    Cc = [rng.normal(size=(C, 3)) for _ in range(N)]
    p = clipio.save_camera_npz(tmp_path / "cams.pt", K, R, t, Cc)     # vggt/save.py:108: with_suffix(".npz")
    assert p.name == "cams.npz"
    d = clipio.load_camera_npz(p)
    assert sorted(d) == ["C", "R", "camera_intrinsics", "t"] and d["R"].shape == (N, C, 3, 3)
    np.testing.assert_array_equal(d["camera_intrinsics"], np.stack(K))
    np.testing.assert_allclose(clipio.mean_intrinsics(d["camera_intrinsics"]), np.mean(np.stack(K), 0))


def test_stereo_clip_staging_cpu():
    xy = np.random.default_rng(1).uniform(0, 1000, (5, 17, 2)).astype(np.float32)
    sc = np.ones((5, 17), np.float32)
    k, c = clipio.stereo_clip_to_device((xy, sc), (xy[:4] + 1, sc[:4]), device="cpu")
    assert k.shape == (2, 4, 17, 2) and c.shape == (2, 4, 17)
    np.testing.assert_array_equal(k[1].numpy(), xy[:4] + 1)
    with pytest.raises(ValueError):
        clipio.stereo_clip_to_device((xy, sc), (xy[:, :16], sc[:, :16]), device="cpu")


@pytest.mark.gpu
def test_files_to_gpu_pipeline(cuda, golden):
    """file -> pinned host -> GPU -> batch API for both paths the loaders feed."""
    from skiing_analysis_pytorch_b200 import api, fusion, synth

    k, c = clipio.stereo_clip_to_device(clipio.load_keypoints_pt(IO / "left.pt"), clipio.load_keypoints_pt(IO / "right.pt"), cuda)
    R, t = synth.rig("2b")
    res = api.triangulate_reproject(k, synth.K_CALIB, R, t, conf=c, want=("X", "err"))
    assert res.X.shape == (9, 17, 3) and res.err.shape == (2, 9, 17)
    Xl, Xr, Ul, Ur = (torch.from_numpy(a).to(cuda) for a in clipio.load_sam3d_pair({"sam_l": IO / "osmo_2_sam_3d_body_outputs.npz", "sam_r": IO / "right"}))
    r = fusion.fuse_clip(Xl, Xr, Ul, Ur)
    assert r.fused.shape == (6, 70, 3) and int(r.status.sum()) == 0
