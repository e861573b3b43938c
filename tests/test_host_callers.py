"""CPU tests of the host-side helpers around the step path that need no GPU: the export module's geometry against the
reference's own pickles, record classes, rollout-buffer contract checks."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "episode_storage_obstacles.npz")


def test_export_end_points_match_reference_radar_rows():
    """Radar.end_pos_n = pos + normalise(Rzyx(Theta) rd_b) * dist (sensor.py:96-120): with every ray at max_dist the
    reference's logged end points must equal the exporter's geometry for the logged state (rows where nothing was hit)."""
    from gym_dockauv_b200.config import BASE_CONFIG
    from gym_dockauv_b200.episode_export import _rzyx
    from gym_dockauv_b200.params import radar_geometry
    g = np.load(GOLD)
    r = BASE_CONFIG["radar"]
    geo = radar_geometry(r["alpha"], r["beta"], r["ray_per_deg"], r["max_dist"], r["blocksize_reduce"])
    rd_b, dmax = geo["rd_b"], geo["max_dist"]
    checked = 0
    for e in g["episodes"]:
        states, radar = g[f"states_{e}"], g[f"radar_{e}"]
        for t in range(1, len(states) - 1):
            s = states[t]
            rd_n = rd_b.dot(_rzyx(*s[3:6]).T)
            rd_n /= np.linalg.norm(rd_n, axis=1)[:, None]
            full = s[0:3] + rd_n * dmax
            miss = np.linalg.norm(radar[t] - s[0:3], axis=1) > dmax * (1 - 1e-12)      # rays the reference left at max_dist
            assert miss.any()
            assert np.max(np.abs(full[miss] - radar[t][miss])) < 1e-9
            checked += int(miss.sum())
    assert checked > 10000
    # row 0 is logged by the radar reset at the origin with zero attitude (docking3d.py:262,285)
    assert np.max(np.abs(g["radar_0"][0] - rd_b / np.linalg.norm(rd_b, axis=1)[:, None] * dmax)) < 1e-12


def test_rzyx_is_a_rotation_and_matches_known_answers():
    from gym_dockauv_b200.episode_export import _rzyx
    assert np.array_equal(_rzyx(0.0, 0.0, 0.0), np.identity(3))
    R = _rzyx(0.2, -0.3, 1.0)
    assert np.allclose(R.dot(R.T), np.identity(3), atol=1e-15) and abs(np.linalg.det(R) - 1) < 1e-15
    # pure yaw by 90 deg maps x_b to y_n (geomutils.py:40-43)
    assert np.allclose(_rzyx(0, 0, np.pi / 2).dot([1, 0, 0]), [0, 1, 0], atol=1e-15)


def test_shape_record_mirrors_reference_capsule_geometry():
    from gym_dockauv_b200.episode_export import META_DATA_REWARD, N_CONT_REWARDS, ShapeRecord, meta_data_observation
    c = ShapeRecord("capsule", [1.0, 2.0, 0.0], 1.0, vec_top=[1.0, 2.0, -20.0])
    assert np.array_equal(c.vec_bot, [1.0, 2.0, 20.0])            # vec_bot = 2 * position - vec_top, shape.py:105-108
    assert len(META_DATA_REWARD) == 13 and META_DATA_REWARD[N_CONT_REWARDS:] == [
        "Done-Goal_reached", "Done-out_pos", "Done-out_att", "Done-max_t", "Done-collision"]
    md = meta_data_observation(20)
    assert sum(len(m) for m in md) == 36 and md[-1][0] == "ray_0"


def test_recorder_and_buffer_refuse_wrong_env_modes():
    from gym_dockauv_b200.episode_export import EpisodeRecorder
    from gym_dockauv_b200.reference_reset import ReferenceSeededEnv

    class FakeEnv:
        auto_reset, debug, num_envs = True, None, 2

    with pytest.raises(ValueError):
        EpisodeRecorder(FakeEnv(), [0], "x")
    with pytest.raises(ValueError):
        ReferenceSeededEnv(FakeEnv(), [0, 1])
