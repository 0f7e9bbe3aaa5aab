"""The reference's YAML formats: byte-level writer checks against the literal
`fout <<` statements of corner_detections.cpp:18-39 / camera_pose.cpp:95-129,
reader round trips, and (GPU) the Milestone-3 directory driver."""
import os

import numpy as np
import pytest
import yaml

from robot_camera_calibration_b200 import io_yaml
from robot_camera_calibration_b200.scenes import make_scene


def test_detections_text_is_byte_compatible():
    txt = io_yaml.detections_text([7], [0.1], np.array([[[10, 20], [30, 21], [31, 40], [11, 41]]]))
    # corner_detections.cpp:27,31-37,59
    assert txt == ("detections:"
                   "\n - targetID: 7"
                   "\n   size: [ 0.100000, 0.100000 ]"
                   "\n   corners:"
                   "\n    0: [ 10, 20 ]"
                   "\n    1: [ 30, 21 ]"
                   "\n    2: [ 31, 40 ]"
                   "\n    3: [ 11, 41 ]"
                   "\n")
    y = yaml.safe_load(txt)
    assert y["detections"][0]["corners"][2] == [31, 40]      # yaml-cpp indexes the same node as "2"


def test_world_T_camera_and_targets_text_are_byte_compatible():
    w = io_yaml.world_T_camera_text([0.1, -0.2, 0.3], [1.0, 2.5, -3.25])
    # camera_pose.cpp:96-98
    assert w == ("world_T_camera:"
                 "\n rotation: [ 0.100000 , -0.200000 , 0.300000 ]"
                 "\n translation: [ 1.000000 , 2.500000 , -3.250000 ]")
    t = io_yaml.targets_text([3], [0.2], [[0, 0, 0, 0, 0, 0]])
    # camera_pose.cpp:107,118-126
    assert t == ("targets:"
                 "\n - targetID: 3"
                 "\n   world_T_target:"
                 "\n    rotation: [ 0.000000 , 0.000000 , 0.000000 ]"
                 "\n    translation: [ 0.000000 , 0.000000 , 0.000000 ]"
                 "\n   obj_points_in_target:"
                 "\n    0: [ -0.100000, -0.100000, 0 ]"
                 "\n    1: [ 0.100000, -0.100000, 0 ]"
                 "\n    2: [ 0.100000, 0.100000, 0 ]"
                 "\n    3: [ -0.100000, 0.100000, 0 ]")
    y = yaml.safe_load(t)["targets"][0]
    assert 2.0 * y["obj_points_in_target"][2][0] == pytest.approx(0.2)   # opt_visualization.cpp:81


def test_dataset_round_trip(tmp_path):
    s = make_scene(9, 11, 0.8, seed=3)
    ids = np.arange(9) * 3 + 5                       # arbitrary AprilTag ids
    io_yaml.write_dataset(str(tmp_path), s, tag_ids=ids)
    assert len(list(tmp_path.glob("detections_*.yaml"))) == 11
    r, tag_ids, frames = io_yaml.read_dataset(str(tmp_path))
    assert list(tag_ids) == list(ids) and list(frames) == list(range(11))
    assert r.n_blocks == s.n_blocks
    assert np.abs(r.views - s.views).max() < 1e-6      # 6-decimal text (camera_pose.cpp:97)
    assert np.abs(r.markers - s.markers).max() < 1e-6
    assert np.allclose(r.sizes, s.sizes, atol=2e-6)
    assert np.allclose(r.intr, s.intr) and np.allclose(r.dist, s.dist)
    # observations: same multiset of (view, marker, truncated pixels)
    key = lambda sc: sorted(zip(sc.view_idx.tolist(), sc.marker_idx.tolist(), map(tuple, np.trunc(sc.pixels).tolist())))
    assert key(r) == key(s)
    # frame 0 lists the world tag first (camera_pose.cpp:74)
    y0 = yaml.safe_load(open(tmp_path / "detections_0.yaml"))
    if 0 in s.marker_idx[s.view_idx == 0]:
        assert y0["detections"][0]["targetID"] == ids[0]


def test_frames_without_pose_and_unknown_tags_are_skipped(tmp_path):
    s = make_scene(6, 5, 0.9, seed=4)
    io_yaml.write_dataset(str(tmp_path), s)
    # frame 2 was never referenced by camera_pose_node: no world_T_camera stanza
    p = tmp_path / "detections_2.yaml"
    txt = p.read_text()
    p.write_text(txt[:txt.find("world_T_camera:")])
    # a tag that is not in targets.yaml
    p = tmp_path / "detections_1.yaml"
    txt = p.read_text()
    p.write_text(txt.replace("detections:", "detections:\n - targetID: 999\n   size: [ 0.1, 0.1 ]\n   corners:\n"
                                            "    0: [ 1, 1 ]\n    1: [ 2, 1 ]\n    2: [ 2, 2 ]\n    3: [ 1, 2 ]", 1))
    r, tag_ids, frames = io_yaml.read_dataset(str(tmp_path))
    assert list(frames) == [0, 1, 3, 4]
    assert r.n_blocks == int((s.view_idx != 2).sum())


def test_write_results_replaces_the_stanza(tmp_path):
    s = make_scene(5, 4, 0.9, seed=6)
    io_yaml.write_dataset(str(tmp_path), s)
    r, ids, frames = io_yaml.read_dataset(str(tmp_path))
    r.views[:, 3:6] += 1.0
    r.markers[1:, 0:3] *= 0.5
    io_yaml.write_results(str(tmp_path), r, ids, frames, precision="repr")
    r2, _, _ = io_yaml.read_dataset(str(tmp_path))
    assert np.array_equal(r2.views, r.views) and np.array_equal(r2.markers, r.markers)
    assert r2.n_blocks == r.n_blocks
    assert (tmp_path / "detections_0.yaml").read_text().count("world_T_camera:") == 1


@pytest.mark.gpu
def test_optimise_directory_is_milestone_3(tmp_path):
    s = make_scene(10, 25, 0.8, seed=8, pixel_noise=0.0, perturb=(0.01, 0.01, 0.0))
    io_yaml.write_dataset(str(tmp_path), s)
    before, _, _ = io_yaml.read_dataset(str(tmp_path))
    summ = io_yaml.optimise_directory(str(tmp_path), refine_intrinsics=False, precision="repr", max_iterations=40)
    after, _, _ = io_yaml.read_dataset(str(tmp_path))
    assert summ["final_cost"] < 0.2 * summ["initial_cost"]
    # integer-truncated pixels limit the accuracy; the refined map must still be closer to truth
    assert np.abs(after.markers - s.truth["markers"]).mean() < np.abs(before.markers - s.truth["markers"]).mean()


def test_missing_intrinsics_or_targets_are_errors_and_integer_corners_stay_integers(tmp_path):
    """A dataset without camera.yaml must not be optimised from made-up intrinsics (the reference logs
    ROS_ERROR, camera_pose.cpp:66-67); an empty targets.yaml has no world frame; the reference's integer corners
    (corner_detections.cpp:53-54) are read back as int16 so that they cross PCIe at 16 bytes per tag."""
    s = make_scene(6, 5, 0.9, seed=91, round_pixels=True)
    io_yaml.write_dataset(str(tmp_path), s)
    r, _, _ = io_yaml.read_dataset(str(tmp_path))
    assert r.pixels.dtype == np.int16 and np.array_equal(r.pixels, np.trunc(s.pixels).astype(np.int16))
    os.remove(os.path.join(str(tmp_path), "camera.yaml"))
    with pytest.raises(FileNotFoundError):
        io_yaml.read_dataset(str(tmp_path))
    r2, _, _ = io_yaml.read_dataset(str(tmp_path), intr=s.intr[0], dist=s.dist[0])      # explicit intrinsics are fine
    assert np.array_equal(r2.intr[0], s.intr[0])
    with open(os.path.join(str(tmp_path), "targets.yaml"), "w") as f:
        f.write("targets:\n")
    with pytest.raises(ValueError):
        io_yaml.read_dataset(str(tmp_path), intr=s.intr[0], dist=s.dist[0])


def test_command_line_help_and_dataset_errors(tmp_path, capsys):
    """python -m robot_camera_calibration_b200 <detections directory>: argument parsing, and a dataset without
    camera.yaml is refused (exit code 2) before any device is touched."""
    from robot_camera_calibration_b200.__main__ import main, parser
    a = parser().parse_args([str(tmp_path), "--fix-intrinsics", "--max-iterations", "7"])
    assert a.fix_intrinsics and a.max_iterations == 7 and a.device == 0
    s = make_scene(6, 5, 0.9, seed=92, round_pixels=True)
    io_yaml.write_dataset(str(tmp_path), s)
    os.remove(os.path.join(str(tmp_path), "camera.yaml"))
    assert main([str(tmp_path)]) == 2
    assert "camera.yaml" in capsys.readouterr().err


@pytest.mark.gpu
def test_command_line_runs_milestone_3(tmp_path, capsys):
    import json
    from robot_camera_calibration_b200.__main__ import main
    s = make_scene(10, 25, 0.8, seed=8, pixel_noise=0.0, perturb=(0.01, 0.01, 0.0))
    io_yaml.write_dataset(str(tmp_path), s)
    before, _, _ = io_yaml.read_dataset(str(tmp_path))
    assert main([str(tmp_path), "--fix-intrinsics", "--precision", "repr", "--max-iterations", "40"]) == 0
    summ = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert summ["final_cost"] < 0.2 * summ["initial_cost"] and summ["termination_name"] != "failure"
    after, _, _ = io_yaml.read_dataset(str(tmp_path))
    assert np.abs(after.markers - s.truth["markers"]).mean() < np.abs(before.markers - s.truth["markers"]).mean()
