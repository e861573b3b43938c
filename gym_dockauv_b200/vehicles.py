"""Vehicle parameter tables (BlueROV2, LAUV) and the loader for the reference's flat XML parameter files.

The numbers are the physical constants of the reference's vehicle files (objects/vehicles/BlueROV2.xml,
LAUV.xml, BlueROV2.py:34-72, LAUV.py:103-110); tests/test_host_params.py checks them against the values the
reference's own objects hold (tests/golden/vehicles.json).
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

# every hydrodynamic derivative the StateSpace base knows (statespace.py:54-83); missing entries are 0
_BASE_KEYS = ("m", "BY", "I_x", "I_y", "I_z", "I_xy", "I_xz", "I_yz", "x_G", "y_G", "z_G", "x_B", "y_B", "z_B",
              "X_udot", "Y_vdot", "Z_wdot", "K_pdot", "M_qdot", "N_rdot", "X_u", "Y_v", "Z_w", "K_p", "M_q", "N_r",
              "X_uu", "Y_vv", "Z_ww", "K_pp", "M_qq", "N_rr")
# the extra coefficients of the LAUV model (LAUV.py:32-55)
_LAUV_KEYS = ("N_urf", "N_uvf", "N_uvb", "M_uqf", "M_uwf", "M_uwb", "Z_uqf", "Z_uwf", "Z_uwb", "Y_urf", "Y_uvf",
              "Y_uvb", "N_vv", "M_ww", "Z_qq", "Y_rr", "N_v", "M_w", "Z_q", "Y_r", "N_uudr", "M_uuds", "Z_uuds",
              "Y_uudr")

GRAVITY = 9.81          # statespace.py:62
LOWPASS_T1 = 0.2        # auvsim.py:40
SAFETY_RADIUS = 1.0     # auvsim.py:43 (the config key "radius" is never read by the reference)


def _table(pairs):
    d = {k: 0.0 for k in _BASE_KEYS}
    d.update(pairs)
    return d


_BLUEROV2 = _table(dict(
    m=11.5, BY=114.8, I_x=0.21, I_y=0.245, I_z=0.245, z_G=0.02,
    X_udot=-7.57, Y_vdot=-7.57, Z_wdot=-7.57, K_pdot=-0.12, M_qdot=-0.12, N_rdot=-0.12,
    X_u=-4.03, Y_v=-6.22, Z_w=-5.18, K_p=-0.07, M_q=-0.07, N_r=-0.07,
    X_uu=-18.18, Y_vv=-21.66, Z_ww=-36.99, K_pp=-1.55, M_qq=-1.55, N_rr=-1.55))

_LAUV = _table(dict(
    m=18.0, BY=177.58, I_x=0.0405, I_y=1.07, I_z=1.07, z_G=0.01,
    X_udot=-1.0291, Y_vdot=-16.153, Z_wdot=-16.153, K_pdot=0.0, M_qdot=0.758, N_rdot=0.758,
    X_u=-2.4, Y_v=-23.0, Z_w=-23.0, K_p=-0.3, M_q=-9.7, N_r=-9.7,
    X_uu=-2.4, Y_vv=-80.0, Z_ww=-80.0, K_pp=-0.0006, M_qq=-9.1, N_rr=-9.1,
    N_urf=-3.072, N_uvf=7.68, N_uvb=3.3088, M_uqf=-3.072, M_uwf=-7.68, M_uwb=-3.3088,
    Z_uqf=-7.68, Z_uwf=-19.2, Z_uwb=-10.956, Y_urf=7.68, Y_uvf=-19.2, Y_uvb=-10.956,
    N_vv=-1.5, M_ww=1.5, Z_qq=-0.3, Y_rr=0.3, N_v=-3.1, M_w=3.1, Z_q=-11.5, Y_r=11.5,
    N_uudr=-7.68, M_uuds=-7.68, Z_uuds=-19.2, Y_uudr=19.2))

VEHICLE_NAMES = ("BlueROV2", "LAUV")


def read_phys_para_from_xml(xml_path, vehicle):
    """Flat <Parameter><tag>value</tag>...</Parameter> file as used by the reference (statespace.py:428-448).
    Unknown tags raise AttributeError like the reference does."""
    allowed = set(_BASE_KEYS) | {"name", "version"}
    if vehicle == "LAUV":
        allowed |= set(_LAUV_KEYS)
    out = {}
    for child in ET.parse(xml_path).getroot():
        if child.tag not in allowed:
            raise AttributeError("Bad and not allowed practice: Trying to parse xml data tag without it being "
                                 "initialized in init")
        if child.tag not in ("name", "version"):
            out[child.tag] = float(child.text)
    return out


def bluerov2_input_map(control_mode="joystick"):
    """B (6 x n_u) and u_bound (n_u x 2) of BlueROV2.py:27-72."""
    if control_mode == "joystick":
        B = np.diag([2.83, 2.83, 4.0, 0.436, 0.24, 0.378]) * 20
        n_u = 6
    elif control_mode == "direct":
        T_thrust = np.array([
            [0.707, 0.707, -0.707, -0.707, 0, 0, 0, 0],
            [-0.707, 0.707, -0.707, 0.707, 0, 0, 0, 0],
            [0, 0, 0, 0, -1, -1, -1, -1],
            [0.06, -0.06, 0.06, -0.06, -0.218, -0.218, 0.218, 0.218],
            [0.06, 0.06, -0.06, -0.06, 0.120, -0.120, 0.120, -0.120],
            [-0.189, 0.189, 0.189, -0.189, 0, 0, 0, 0]])
        B = np.dot(T_thrust, np.diag([40.0] * 8))
        n_u = 8
    else:
        raise KeyError("Invalid control mode for BlueROV2 initialization.")
    u_bound = np.tile(np.array([-1.0, 1.0]), (n_u, 1))
    return B, u_bound


def lauv_u_bound():
    """LAUV.py:103-110: thrust [0, 14] N, rudder / stern plane +-30 deg."""
    a = 30 * np.pi / 180
    return np.array([[0.0, 14.0], [-a, a], [-a, a]])


def load_vehicle(name, xml_path=None, control_mode="joystick"):
    """Returns the parameter table of a vehicle ({tag: float}), plus 'B' / 'u_bound' / 'n_u'."""
    if name not in VEHICLE_NAMES:
        raise ModuleNotFoundError(f"No vehicle named {name!r}; available: {VEHICLE_NAMES}")
    table = dict(_BLUEROV2 if name == "BlueROV2" else _LAUV)
    if name == "LAUV":
        for k in _LAUV_KEYS:
            table.setdefault(k, 0.0)
    if xml_path is not None:
        table.update(read_phys_para_from_xml(xml_path, name))
    if name == "BlueROV2":
        B, u_bound = bluerov2_input_map(control_mode)
        table["B"] = B
    else:
        u_bound = lauv_u_bound()
    table["u_bound"] = u_bound
    table["n_u"] = int(u_bound.shape[0])
    table["name"] = name
    return table
