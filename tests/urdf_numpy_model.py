"""Independent numpy model of the Go1 built straight from tests/golden/go1_urdf.json
(the reference URDF parsed by oracle/gen_golden.py).  Used to cross-check the
oracle's and the CUDA kernel's mass matrix / gravity terms.  Pure geometry:
M = sum_links  m Jv^T Jv + Jw^T (R I R^T) Jw  with geometric Jacobians."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LINK_ORDER = ["base", "trunk", "imu_link"] + [f"{leg}_{p}" for leg in ("FR", "FL", "RR", "RL")
                                               for p in ("hip", "thigh", "calf", "foot")]
MOTOR_JOINTS = [f"{leg}_{j}_joint" for leg in ("FR", "FL", "RR", "RL") for j in ("hip", "thigh", "calf")]


def load():
    with open(os.path.join(GOLDEN, "go1_urdf.json")) as f:
        return json.load(f)


def rot_axis(a, th):
    a = np.asarray(a, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def rpy_R(rpy):
    r, p, y = rpy
    return rot_axis([0, 0, 1], y) @ rot_axis([0, 1, 0], p) @ rot_axis([1, 0, 0], r)


def bullet_inertia_diag(link):
    """Bullet's rule when URDF_USE_INERTIA_FROM_FILE is absent."""
    m, cols = link["mass"], link["collision"]
    if not cols:
        return np.zeros(3)
    c = cols[0]
    if c["type"] == "sphere":
        return np.full(3, 0.4 * m * c["radius"][0] ** 2)
    R = np.abs(rpy_R(c["rpy"]))
    if c["type"] == "box":
        ext = R @ np.asarray(c["size"])
    else:
        ext = R @ np.array([2 * c["radius"][0], 2 * c["radius"][0], c["length"][0]])
    lx, ly, lz = ext
    return m / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])


def quat_R(q):
    x, y, z, w = q
    s = 2.0 / (x * x + y * y + z * z + w * w)
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


class Model:
    def __init__(self):
        u = load()
        self.links = u["links"]
        self.joint_of_child = {j["child"]: j for j in u["joints"]}

    def fk(self, pos, quat, q):
        """world pose of every link frame + list of (joint name, world axis, world origin) ancestors"""
        poses = {"base": (quat_R(quat), np.asarray(pos, float))}
        anc = {"base": []}
        for name in LINK_ORDER[1:]:
            j = self.joint_of_child[name]
            Rp, pp = poses[j["parent"]]
            R = Rp @ rpy_R(j["rpy"])
            p = pp + Rp @ np.asarray(j["xyz"])
            a = list(anc[j["parent"]])
            if j["type"] == "revolute":
                dof = MOTOR_JOINTS.index(j["name"])
                a.append((dof, R @ np.asarray(j["axis"]), p))
                R = R @ rot_axis(j["axis"], q[dof])
            poses[name] = (R, p)
            anc[name] = a
        return poses, anc

    def mass_matrix(self, pos, quat, q):
        """M in coordinates nu = [omega_b, v_b (base-frame coords, base origin), qd]."""
        poses, anc = self.fk(pos, quat, q)
        Rb, pb = poses["base"]
        M = np.zeros((18, 18))
        for name in LINK_ORDER:
            L = self.links[name]
            R, p = poses[name]
            c = p + R @ np.asarray(L["com"])
            Jv = np.zeros((3, 18)); Jw = np.zeros((3, 18))
            for k in range(3):
                e = Rb[:, k]
                Jw[:, k] = e
                Jv[:, k] = np.cross(e, c - pb)
                Jv[:, 3 + k] = e
            for dof, ax, org in anc[name]:
                Jw[:, 6 + dof] = ax
                Jv[:, 6 + dof] = np.cross(ax, c - org)
            Iw = R @ np.diag(bullet_inertia_diag(L)) @ R.T
            M += L["mass"] * Jv.T @ Jv + Jw.T @ Iw @ Jw
        return M

    def potential(self, pos, quat, q, g=9.8):
        poses, _ = self.fk(pos, quat, q)
        V = 0.0
        for name in LINK_ORDER:
            L = self.links[name]
            R, p = poses[name]
            V += L["mass"] * g * (p + R @ np.asarray(L["com"]))[2]
        return V

    def total_mass(self):
        return sum(self.links[n]["mass"] for n in LINK_ORDER)
