"""CPU tests of the host side: vehicle tables, parameter packing, config contract (no GPU, no compute calls)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from gym_dockauv_b200 import config as cfgmod
from gym_dockauv_b200 import params as P
from gym_dockauv_b200 import vehicles as V
from tests.golden_utils import GOLDEN_DIR, load_case, unit_vectors


def _ref_tables():
    with open(os.path.join(GOLDEN_DIR, "vehicles.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name,key,mode", [("BlueROV2", "BlueROV2", "joystick"), ("BlueROV2", "BlueROV2_direct", "direct"),
                                           ("LAUV", "LAUV", "joystick")])
def test_vehicle_tables_match_reference_xml(name, key, mode):
    """The built-in tables carry exactly the numbers the reference's vehicle objects hold after reading their XML
    (objects/vehicles/*.xml; dumped by tests/golden/make_golden.py into vehicles.json)."""
    ref = _ref_tables()[key]
    mine = V.load_vehicle(name, control_mode=mode)
    for k, v in ref.items():
        if k in ("B", "u_bound"):
            assert np.array_equal(np.asarray(v), np.asarray(mine[k])), k
        elif k in ("g",):
            assert v == V.GRAVITY
        elif k == "lowpass_T1":
            assert v == V.LOWPASS_T1
        elif k == "safety_radius":
            assert v == V.SAFETY_RADIUS
        elif k in ("step_size", "version", "K_thrust"):
            continue
        else:
            assert mine[k] == v, (k, mine[k], v)


@pytest.mark.parametrize("vehicle", ["BlueROV2", "LAUV"])
def test_packed_matrices_bit_equal_reference(vehicle):
    uv = unit_vectors()
    cfg = dict(cfgmod.BASE_CONFIG, vehicle=vehicle)
    p, meta = P.pack_params(cfg, "SimpleDocking3d")
    assert np.array_equal(np.ctypeslib.as_array(p.M_inv).reshape(6, 6), uv[f"{vehicle}_M_inv"])
    assert np.array_equal(np.ctypeslib.as_array(p.I_b).reshape(3, 3), uv[f"{vehicle}_I_b"])
    assert np.array_equal(meta["rigid_body"]["M_RB"], uv[f"{vehicle}_M_RB"])
    assert np.array_equal(np.diag(np.ctypeslib.as_array(p.MA_diag)), np.abs(uv[f"{vehicle}_M_A"]) * np.sign(uv[f"{vehicle}_M_A"] + 0.0))
    # D(nu) rebuilt from the packed coefficient vectors equals the reference's matrix
    nu = uv[f"{vehicle}_nu"]
    lin, quad, lift = (np.ctypeslib.as_array(x) for x in (p.D_lin, p.D_quad, p.D_lift))
    D = np.zeros((6, 6))
    pos = [(i, i) for i in range(6)] + [(1, 5), (2, 4), (4, 2), (5, 1)]
    for k, (i, j) in enumerate(pos):
        D[i, j] = -(lin[k] + quad[k] * abs(nu[j]) + lift[k] * abs(nu[0]))
    assert np.allclose(D, uv[f"{vehicle}_D"], rtol=1e-14, atol=1e-14)
    assert np.array_equal(meta["u_bound"], uv[f"{vehicle}_u_bound"])


@pytest.mark.parametrize("tag,radar", [("stock", None), ("r64", cfgmod.RADAR_64)])
def test_radar_table_matches_reference(tag, radar):
    uv = unit_vectors()
    cfg = dict(cfgmod.BASE_CONFIG)
    if radar is not None:
        cfg["radar"] = dict(radar)
    p, meta = P.pack_params(cfg, "ObstaclesDocking3d")
    n_v, n_h, n_red = uv[f"radar_{tag}_shape"]
    assert (p.n_vert, p.n_horiz, meta["n_obs"] - 16) == (n_v, n_h, n_red)
    assert np.array_equal(meta["radar"]["rd_b"], uv[f"radar_{tag}_rd_b"])
    assert np.array_equal(meta["radar"]["alpha"], uv[f"radar_{tag}_alpha"])
    assert np.array_equal(meta["radar"]["beta"], uv[f"radar_{tag}_beta"])
    assert np.array_equal(np.ctypeslib.as_array(p.rd_b)[:3 * p.n_rays].reshape(-1, 3), uv[f"radar_{tag}_rd_b"])


def test_config_contract():
    """Same keys / defaults as the reference's BASE_CONFIG (config/env_config.py:20-91), taken from a golden trace."""
    ref = load_case("simple_bluerov2_f64")["meta"]["config"]
    mine = cfgmod.BASE_CONFIG
    for k, v in ref.items():
        if k in ("verbose", "log_level", "interval_datastorage", "interval_episode_log"):
            continue   # overridden by the recorder
        assert k in mine, k
        if isinstance(v, dict):
            for kk, vv in v.items():
                assert mine[k][kk] == pytest.approx(vv, rel=0, abs=0), (k, kk)
        else:
            assert mine[k] == v, k
    assert set(cfgmod.REGISTRATION_DICT) == {f"{n}-v0" for n in P.SCENARIO_IDS}
    for cfg in (cfgmod.TRAIN_CONFIG, cfgmod.PREDICT_CONFIG, cfgmod.MANUAL_CONFIG):
        assert cfg["t_step_size"] == 0.10 and cfg["radar"]["max_dist"] == 10
    assert cfgmod.PREDICT_CONFIG["interval_datastorage"] == 1


def test_config_errors_match_reference_behaviour():
    bad = dict(cfgmod.BASE_CONFIG)
    del bad["max_timesteps"]
    with pytest.raises(KeyError):
        P.pack_params(bad, "SimpleDocking3d")
    with pytest.raises(ModuleNotFoundError):      # the reference does importlib.import_module on the vehicle name
        P.pack_params(dict(cfgmod.BASE_CONFIG, vehicle="Nautilus"), "SimpleDocking3d")
    with pytest.raises(KeyError):                 # sensor.py:49-50
        P.pack_params(dict(cfgmod.BASE_CONFIG, radar=dict(cfgmod.BASE_CONFIG["radar"], ray_per_deg=0.3)), "SimpleDocking3d")
    with pytest.raises(KeyError):
        V.bluerov2_input_map("warp")


def test_scenario_capsule_counts_and_sizes():
    for name, n_caps in [("SimpleDocking3d", 0), ("CapsuleDocking3d", 1), ("ObstaclesDocking3d", 5),
                         ("ObstaclesNoCapDocking3d", 4), ("ObstaclesCurrentDocking3d", 5)]:
        p, meta = P.pack_params(cfgmod.BASE_CONFIG, name)
        assert p.n_capsules == n_caps and meta["n_obs"] == 36 and p.n_rays == 63
    p, meta = P.pack_params(dict(cfgmod.BASE_CONFIG, radar=dict(cfgmod.RADAR_64)), "ObstaclesDocking3d",
                            n_synthetic_spheres=3)
    assert (p.n_rays, meta["n_obs"], p.n_spheres) == (64, 32, 3)
    p, meta = P.pack_params(dict(cfgmod.BASE_CONFIG, vehicle="LAUV"), "CapsuleCurrentDocking3d")
    assert p.n_u == 3 and p.vehicle == 1 and meta["u_bound"][0].tolist() == [0.0, 14.0]
    assert p.lp_alpha == pytest.approx(0.1 / 0.3)            # h / (h + T1) = 1/3 at h = 0.1


def test_xml_loader(tmp_path):
    xml = tmp_path / "veh.xml"
    xml.write_text("<Parameter><name>BlueROV2</name><version>1.0</version><m>12.5</m><X_udot>-5.5</X_udot></Parameter>")
    t = V.load_vehicle("BlueROV2", xml_path=str(xml))
    assert t["m"] == 12.5 and t["X_udot"] == -5.5 and t["Y_vdot"] == -7.57
    xml.write_text("<Parameter><bogus>1</bogus></Parameter>")
    with pytest.raises(AttributeError):
        V.load_vehicle("BlueROV2", xml_path=str(xml))


def test_shard_ranges():
    from gym_dockauv_b200.stats import shard_env_range
    for n, w in [(1 << 20, 8), (1000003, 8), (7, 8), (64, 1)]:
        spans = [shard_env_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
