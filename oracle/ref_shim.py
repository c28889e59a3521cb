"""Run the UNMODIFIED reference package in this container (test infrastructure).

The reference (`/root/reference/quadruped_spring`) cannot be imported as
shipped: it needs `gym`, `pybullet`, `pybullet_data`, `pybullet_utils`,
`absl`, `stable_baselines3`, `matplotlib`, and `collections.Sequence`
(removed in Python 3.10, used at env/quadruped_motor.py:38).  None of those are
installable here (no network).  This module installs in-memory stand-ins so the
reference's OWN numpy code runs unchanged:

* every analytic module (springs, motor model, control interfaces, tasks,
  sensors, action filter, HopfNetwork) runs exactly as written;
* `pybullet_utils.bullet_client.BulletClient` is replaced by `FakeBulletClient`,
  which implements the ~35 API calls the reference makes on top of the CPU
  oracle's physics world (oracle/qso_physics.c).  The reference's
  QuadrupedGymEnv.reset()/step() therefore execute their real control flow;
  only `stepSimulation` and the state queries come from our restatement of
  Bullet ("parity unpinned" for that layer, see oracle/qso.h).

It is used ONLY by oracle/gen_golden.py (run here, never on the GPU box) to
produce the fixtures under tests/golden/.
"""
import collections
import collections.abc
import math
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


# --------------------------------------------------------------------------- math helpers
def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    )


def _quat_rot(q, v):
    x, y, z, w = q
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    R = np.array(
        [
            [1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
            [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
            [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)],
        ]
    )
    return R @ np.asarray(v, dtype=float), R


def invert_transform(position, orientation):
    x, y, z, w = orientation
    qi = (-x, -y, -z, w)
    p, _ = _quat_rot(qi, -np.asarray(position, dtype=float))
    return tuple(p), qi


# --------------------------------------------------------------------------- fake pybullet
_JOINT_NAMES = ["floating_base", "imu_joint"] + [
    f"{leg}_{j}" for leg in ("FR", "FL", "RR", "RL") for j in ("hip_joint", "thigh_joint", "calf_joint", "foot_fixed")
]


class FakeBulletClient:
    """Subset of pybullet's API used by the reference, over the oracle world."""

    GUI = 1
    DIRECT = 2
    TORQUE_CONTROL = 2
    VELOCITY_CONTROL = 0
    POSITION_CONTROL = 1
    URDF_USE_SELF_COLLISION = 8
    JOINT_FIXED = 4
    LINK_FRAME = 1
    GEOM_BOX = 3
    COV_ENABLE_PLANAR_REFLECTION = 0
    COV_ENABLE_RGB_BUFFER_PREVIEW = 0
    COV_ENABLE_DEPTH_BUFFER_PREVIEW = 0
    COV_ENABLE_SEGMENTATION_MARK_PREVIEW = 0
    COV_ENABLE_GUI = 0

    world_params = {}  # class-level overrides applied to every new world

    def __init__(self, connection_mode=None, options=""):
        from oracle.oracle import World

        self._World = World
        self.world = None
        self._plane = 0
        self._robot = 1
        self._num_iters = 50
        self._dt = 1.0 / 240.0
        self._gravity = 0.0
        self._mu_ground = 1.0  # plane.urdf has no friction tag: Bullet default 0.5 until changeDynamics
        self._mu_link = 0.5
        self.calls = collections.Counter()

    # world management
    def resetSimulation(self):
        self.world = None
        self._num_iters = 50
        self._dt = 1.0 / 240.0
        self._gravity = 0.0
        self._mu_ground = 0.5
        self._mu_link = 0.5

    def setPhysicsEngineParameter(self, numSolverIterations=None, **kw):
        if numSolverIterations is not None:
            self._num_iters = int(numSolverIterations)
        self._push()

    def setTimeStep(self, dt):
        self._dt = dt
        self._push()

    def setGravity(self, x, y, z):
        self._gravity = z
        self._push()

    def _push(self):
        if self.world is not None:
            self.world.set_params(
                dt=self._dt, num_iterations=self._num_iters, gravity_z=self._gravity,
                mu_ground=self._mu_ground, mu_link=self._mu_link, **self.world_params,
            )

    def loadURDF(self, path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), flags=0, **kw):
        if path.endswith("plane.urdf"):
            return self._plane
        assert path.endswith("go1.urdf"), path
        self.world = self._World()
        s = self.world.get_state()
        s[:] = 0
        s[0:3] = basePosition
        s[3:7] = baseOrientation
        self.world.set_state(s)
        self._push()
        return self._robot

    def changeVisualShape(self, *a, **k):
        pass

    def configureDebugVisualizer(self, *a, **k):
        pass

    def disconnect(self):
        self.world = None

    # model queries
    def getNumJoints(self, body):
        return 18

    def getJointInfo(self, body, i):
        return (i, _JOINT_NAMES[i].encode("UTF-8"))

    def getDynamicsInfo(self, body, link):
        if body == self._block:
            return (self._block_mass, 0.5, (0.0,) * 3, (0.0,) * 3)
        m, I, c = self.world.dynamics(link)
        return (m, self._mu_link, tuple(I), tuple(c))

    def changeDynamics(self, body, link, mass=None, lateralFriction=None, linearDamping=None,
                       angularDamping=None, maxJointVelocity=None, **kw):
        if lateralFriction is not None:
            if body == self._plane:
                self._mu_ground = lateralFriction
            elif body == self._robot:
                self._mu_link = lateralFriction  # the reference sets all links to the same value
            self._push()
        if maxJointVelocity is not None and body == self._robot:
            self.world.set_params(max_coord_vel=maxJointVelocity)
        if mass is not None and body == self._robot:
            self.world.L.qso_world_set_mass(self.world.h, link, float(mass))
        # linear/angular damping: the oracle world has none (quadruped.py:663-668 zeroes them)

    # payload block of the mass randomizer (quadruped.py:778-819): a second body held by a JOINT_FIXED constraint
    GEOM_BOX = 3
    JOINT_FIXED = 4
    _block = 2

    def createCollisionShape(self, shapeType, halfExtents=None, collisionFramePosition=None, **kw):
        return 0

    def createMultiBody(self, baseMass=0, baseCollisionShapeIndex=-1, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), **kw):
        self._block_mass = float(baseMass)
        return self._block

    def createConstraint(self, parentBodyUniqueId, parentLinkIndex, childBodyUniqueId, childLinkIndex, jointType, jointAxis,
                         parentFramePosition, childFramePosition, **kw):
        assert parentBodyUniqueId == self._robot and parentLinkIndex == -1 and childBodyUniqueId == self._block
        assert jointType == self.JOINT_FIXED and not np.any(np.asarray(parentFramePosition))
        # the block's point childFramePosition sits on the base origin: block centre = base origin - childFramePosition
        self._block_delta = -np.asarray(childFramePosition, dtype=float)
        self.world.set_payload(self._block_mass, self._block_delta)
        return 0

    def setCollisionFilterPair(self, *a, **k):
        pass

    # state
    def resetBasePositionAndOrientation(self, body, pos, orn):
        s = self.world.get_state()
        s[0:3] = pos
        s[3:7] = orn
        self.world.set_state(s)

    def resetBaseVelocity(self, body, lin, ang):
        s = self.world.get_state()
        s[7:10] = lin
        s[10:13] = ang
        self.world.set_state(s)

    def resetJointState(self, body, joint, angle, targetVelocity=0):
        dof = self._dof(joint)
        s = self.world.get_state()
        s[13 + dof] = angle
        s[25 + dof] = targetVelocity
        self.world.set_state(s)

    @staticmethod
    def _dof(joint):
        leg, j = divmod(joint - 2, 4)
        assert 0 <= leg < 4 and j < 3, joint
        return 3 * leg + j

    def getBasePositionAndOrientation(self, body):
        self.calls["getBasePositionAndOrientation"] += 1
        s = self.world.get_state()
        if body == self._block:
            d, _ = _quat_rot(tuple(s[3:7]), tuple(self._block_delta))
            return tuple(np.asarray(s[0:3]) + d), tuple(s[3:7])
        return tuple(s[0:3]), tuple(s[3:7])

    def getBaseVelocity(self, body):
        self.calls["getBaseVelocity"] += 1
        s = self.world.get_state()
        return tuple(s[7:10]), tuple(s[10:13])

    def getJointState(self, body, joint):
        self.calls["getJointState"] += 1
        dof = self._dof(joint)
        s = self.world.get_state()
        return (s[13 + dof], s[25 + dof], (0,) * 6, 0.0)

    def setJointMotorControl2(self, bodyIndex, jointIndex, controlMode, force=0, **kw):
        self.calls["setJointMotorControl2"] += 1
        if controlMode == self.TORQUE_CONTROL:
            tau = np.zeros(12)
            tau[self._dof(jointIndex)] = force
            self.world.add_torque(tau)
        # VELOCITY_CONTROL with force=0 neutralises the default motor (quadruped.py:499-507)

    def stepSimulation(self):
        self.calls["stepSimulation"] += 1
        self.world.step()

    def getContactPoints(self):
        self.calls["getContactPoints"] += 1
        out = []
        for link, nf, dist, pos in self.world.contacts():
            out.append((0, self._robot, self._plane, link, -1, tuple(pos), tuple(pos), (0, 0, 1), dist, nf))
        # URDF_USE_SELF_COLLISION (quadruped.py:530-543): rows with bodyA == bodyB, which GetContactInfo counts as
        # invalid when a calf is involved (quadruped.py:236-241)
        for la, lb, dist in self.world.self_contacts():
            out.append((0, self._robot, self._robot, la, lb, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0, 0, 1), dist, 0.0))
        return out

    # transforms
    def invertTransform(self, position, orientation):
        return invert_transform(position, orientation)

    def multiplyTransforms(self, positionA, orientationA, positionB, orientationB):
        p, _ = _quat_rot(orientationA, positionB)
        return tuple(np.asarray(positionA, dtype=float) + p), _quat_mul(orientationA, orientationB)

    def getQuaternionFromEuler(self, rpy):
        r, p, y = rpy
        cr, sr, cp, sp, cy, sy = math.cos(r / 2), math.sin(r / 2), math.cos(p / 2), math.sin(p / 2), math.cos(y / 2), math.sin(y / 2)
        return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)

    def getEulerFromQuaternion(self, q):
        from oracle.oracle import rpy_from_quat

        return tuple(rpy_from_quat(q))

    def getMatrixFromQuaternion(self, q):
        _, R = _quat_rot(q, (0, 0, 0))
        return tuple(R.reshape(9))

    def resetDebugVisualizerCamera(self, *a, **k):
        pass


# --------------------------------------------------------------------------- stub modules
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = dtype

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class _Env:
    metadata = {}

    @property
    def unwrapped(self):
        return self


class _Wrapper(_Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)


_INSTALLED = False


def install():
    """Idempotently install the stand-in modules and put the reference on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence  # quadruped_motor.py:38
    spaces = _module("gym.spaces", Box=_Box)
    reg = _module("gym.envs.registration", register=lambda **kw: None)
    envs = _module("gym.envs", registration=reg)
    _module("gym", Env=_Env, Wrapper=_Wrapper, spaces=spaces, envs=envs)
    _module("pybullet", invertTransform=invert_transform, GUI=1, DIRECT=2)
    _module("pybullet_data", getDataPath=lambda: "/nonexistent/pybullet_data")
    bc = _module("pybullet_utils.bullet_client", BulletClient=FakeBulletClient)
    _module("pybullet_utils", bullet_client=bc)
    logging = _module("absl.logging", info=lambda *a, **k: None)
    _module("absl", logging=logging)
    env_util = _module("stable_baselines3.common.env_util", is_wrapped=lambda env, cls: False,
                       make_vec_env=None)
    common = _module("stable_baselines3.common", env_util=env_util)
    _module("stable_baselines3", common=common)
    pyplot = _module("matplotlib.pyplot")
    _module("matplotlib", use=lambda *a, **k: None, pyplot=pyplot)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True
