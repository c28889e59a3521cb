"""Robot constants with the reference's names.

Mirror of quadruped_spring/go1/configs_go1_with_springs.py and
configs_go1_without_springs.py (values restated, line numbers in comments) so
code that reaches for `env._robot_config.X` keeps working.  The CUDA library
holds its own copy of the numbers it needs (csrc/qs_model_host.h); the test
suite checks both against tests/golden/analytic.npz.
"""
from types import SimpleNamespace

import numpy as np


def go1_config(enable_springs: bool) -> SimpleNamespace:
    c = SimpleNamespace()
    c.NUM_MOTORS, c.NUM_LEGS, c.MOTORS_PER_LEG = 12, 4, 3                      # :19-21
    c.INIT_RACK_POSITION = [0, 0, 1]
    c.INIT_POSITION = [0, 0, 0.32]                                             # :23
    c.IS_FALLEN_HEIGHT = 0.10 if enable_springs else 0.12                      # :24
    c.INIT_ORIENTATION = (0, 0, 0, 1)
    c.INIT_ORIENTATION_INV = (0, 0, 0, 1)
    c.DEFAULT_HIP_ANGLE, c.DEFAULT_THIGH_ANGLE, c.DEFAULT_CALF_ANGLE = 0, np.pi / 4, -np.pi / 2   # :31-33
    c.INIT_JOINT_ANGLES = np.array([0, np.pi / 4, -np.pi / 2] * 4)
    c.INIT_MOTOR_ANGLES = c.INIT_JOINT_ANGLES
    c.ANGLE_LANDING_POSE = c.INIT_MOTOR_ANGLES
    c.ANGLE_SETTLING_POSE = np.array([0.0, 1.14, -2.5 if enable_springs else -2.19] * 4)
    c.JOINT_DIRECTIONS = np.ones(12)
    c.JOINT_OFFSETS = np.zeros(12)
    c.HIP_LINK_LENGTH, c.THIGH_LINK_LENGTH, c.CALF_LINK_LENGTH = 0.0847, 0.213, 0.213   # :56-58
    c.X_OFFSET, c.Y_OFFSET = 0.1881, 0.04675
    c.DEFAULT_X, c.DEFAULT_Y, c.DEFAULT_Z, c.LANDING_Z = 0, 0.0847, -0.32, -0.29
    sign = np.array([-1, 1, -1, 1])
    c.NOMINAL_FOOT_POS_LEG_FRAME = np.array([[0, s * 0.0847, -0.32] for s in sign]).flatten()
    c.CARTESIAN_LANDING_POSE = np.array([[0, s * 0.0847, -0.29] for s in sign]).flatten()
    c.CARTESIAN_SETTLING_POSE = np.array([[-0.02, s * 0.0847, -0.15] for s in sign]).flatten()
    c.INIT_HEIGHT = 0.35
    c.REAL_UPPER_ANGLE_JOINT = np.array([1.0471975512, 2.96705972839, -0.837758040957] * 4)
    c.REAL_LOWER_ANGLE_JOINT = np.array([-1.0471975512, -0.663225115758, -2.72271363311] * 4)
    c.RL_UPPER_ANGLE_JOINT = np.array([0.2, np.pi / 4 + 0.5, -0.95] * 4)      # :84
    c.RL_LOWER_ANGLE_JOINT = np.array([-0.2, np.pi / 4 - 0.5, -2.5 if enable_springs else -2.12] * 4)
    up_z = 0.18 if enable_springs else 0.11
    c.RL_UPPER_CARTESIAN_POS = c.NOMINAL_FOOT_POS_LEG_FRAME + np.array([0.2, 0.05, up_z] * 4)
    c.RL_LOWER_CARTESIAN_POS = c.NOMINAL_FOOT_POS_LEG_FRAME - np.array([0.2, 0.05, 0.07] * 4)
    c.TORQUE_LIMITS = np.asarray([23.7, 23.7, 33.55] * 4)                      # :100
    c.RL_TORQUE_LIMITS = 1.0 * c.TORQUE_LIMITS
    c.VELOCITY_LIMITS = np.asarray([30.1] * 12)
    c.RL_VELOCITY_LIMITS = np.asarray([10.0] * 12)
    if enable_springs:
        c.MOTOR_KP, c.MOTOR_KD = [75.0, 75.0, 75.0] * 4, [0.8, 1.0, 1.0] * 4   # with:106-107
        c.kpCartesian, c.kdCartesian = np.diag([1200, 2000, 2000]), np.diag([13, 15, 15])
        c.SPRINGS_STIFFNESS = [20, 20, 30]                                     # with:150-160
        c.SPRINGS_DAMPING = [0.3, 0.3, 0.3]
        c.SPRINGS_REST_ANGLE = [0, np.pi / 4, -np.pi / 2 + 0.3]
    else:
        c.MOTOR_KP, c.MOTOR_KD = [55, 60, 60] * 4, [0.8, 1.0, 1.0] * 4         # without:108-109
        c.kpCartesian, c.kdCartesian = np.diag([500, 500, 500]), np.diag([10, 10, 10])
    c.MAX_MOTOR_ANGLE_CHANGE_PER_STEP = 0.2
    c.MAX_CARTESIAN_FOOT_POS_CHANGE_PER_STEP = np.array([0.1, 0.02, 0.08])
    # sensor limits (:176-211)
    c.HEIGHT_HIGH, c.HEIGHT_LOW = np.array([0.4]), np.array([0.1])
    c.VEL_LIN_HIGH = np.array([5.0] * 3); c.VEL_LIN_LOW = -c.VEL_LIN_HIGH
    c.VEL_ANG_HIGH = np.array([3.0] * 3); c.VEL_ANG_LOW = -c.VEL_ANG_HIGH
    c.ORIENT_RPY_HIGH = np.array([np.pi] * 3); c.ORIENT_RPY_LOW = -c.ORIENT_RPY_HIGH
    c.ORIENT_RATE_HIGH = np.array([5.0] * 3)
    c.JOINT_ANGLES_HIGH, c.JOINT_ANGLES_LOW = c.RL_UPPER_ANGLE_JOINT, c.RL_LOWER_ANGLE_JOINT
    c.JOINT_VELOCITIES_HIGH = c.RL_VELOCITY_LIMITS; c.JOINT_VELOCITIES_LOW = -c.JOINT_VELOCITIES_HIGH
    c.CONTACT_BOOL_HIGH, c.CONTACT_BOOL_LOW = np.array([1.0] * 4), np.array([0.0] * 4)
    c.FEET_POS_HIGH, c.FEET_POS_LOW = c.RL_UPPER_CARTESIAN_POS, c.RL_LOWER_CARTESIAN_POS
    c.FEET_VEL_HIGH = np.array([10.0] * 12); c.FEET_VEL_LOW = -c.FEET_POS_HIGH   # sic (:206)
    c.PITCH_HIGH = np.array([np.pi]); c.PITCH_LOW = -c.PITCH_HIGH
    c.PITCH_RATE_HIGH = np.array([5.0]); c.PITCH_RATE_LOW = -c.PITCH_RATE_HIGH
    # sensor noise (:215-230)
    s = 0.01
    c.STD_COEFF = s
    c.HEIGHT_NOISE = c.HEIGHT_HIGH * s * 0.8
    c.VEL_LIN_NOISE = c.VEL_LIN_HIGH * s * 0.8
    c.VEL_ANG_NOISE = c.VEL_ANG_HIGH * s
    c.ORIENT_RPY_NOISE = c.ORIENT_RPY_HIGH * s
    c.ORIENT_RATE_NOISE = c.ORIENT_RATE_HIGH * s
    c.JOINT_ANGLES_NOISE = np.maximum(abs(c.JOINT_ANGLES_HIGH), abs(c.JOINT_ANGLES_LOW)) * s * 0.1
    c.JOINT_VELOCITIES_NOISE = c.JOINT_VELOCITIES_HIGH * s * 0.6
    c.CONTACT_BOOL_NOISE = np.array([0] * 4)
    c.FEET_POS_NOISE = np.array([0.1, 0.05, 0.1] * 4) * s
    c.FEET_VEL_NOISE = c.FEET_VEL_HIGH * s
    c.PITCH_NOISE = c.PITCH_HIGH * s * 0.9
    c.PITCH_RATE_NOISE = c.PITCH_RATE_HIGH * s
    return c


# sensor sets of sensors/sensor_collection.py:18-105 as (name, dim, high, low) builders
def _sensor_table(c):
    one = lambda v: np.atleast_1d(np.asarray(v, dtype=np.float64))
    return {
        "JointPosition": ("Encoder", c.JOINT_ANGLES_HIGH, c.JOINT_ANGLES_LOW),
        "JointVelocity": ("JointVelocity", c.JOINT_VELOCITIES_HIGH, c.JOINT_VELOCITIES_LOW),
        "FeetPostion": ("FeetPosition", c.FEET_POS_HIGH, c.FEET_POS_LOW),
        "FeetVelocity": ("FeetVelocity", c.FEET_VEL_HIGH, c.FEET_VEL_LOW),
        "LinearVelocity": ("Base Linear Velocity", c.VEL_LIN_HIGH, c.VEL_LIN_LOW),
        "AngularVelocity": ("Base Angular Velocity", c.VEL_ANG_HIGH, c.VEL_ANG_LOW),
        "Pitch": ("Pitch", c.PITCH_HIGH, c.PITCH_LOW),
        "PitchRate": ("Pitch rate", c.PITCH_RATE_HIGH, c.PITCH_RATE_LOW),
        "Height": ("Height", c.HEIGHT_HIGH, c.HEIGHT_LOW),
        "BaseHeightVelocity": ("Base Linear Velocity z direction", one(c.VEL_LIN_HIGH[2]), one(c.VEL_LIN_LOW[2])),
        "VelocityX": ("Base Height Velocity X", one(c.VEL_LIN_HIGH[0]), one(c.VEL_LIN_LOW[0])),
        "Landing": ("is landing", one(1), one(0)),
        "Jumping": ("is jumping", one(1), one(0)),
        "BooleanContact": ("BoolContatc", c.CONTACT_BOOL_HIGH, c.CONTACT_BOOL_LOW),
        "PitchBackFlip": ("Pitch-BackFlip", c.PITCH_HIGH, c.PITCH_LOW),
    }


SENSOR_SETS = {
    "ENCODER": ["JointPosition", "JointVelocity"],
    "ENCODER_2": ["LinearVelocity", "AngularVelocity", "JointPosition", "JointVelocity"],
    "CARTESIAN_NO_IMU": ["FeetPostion", "FeetVelocity"],
    "ARS_BASIC": ["JointPosition", "JointVelocity", "Pitch", "Height", "BaseHeightVelocity"],
    "ARS_SENSOR": ["JointPosition", "JointVelocity", "Pitch", "PitchRate", "Height", "BaseHeightVelocity"],
    "LANDING_SENSOR": ["JointPosition", "JointVelocity", "Pitch", "PitchRate", "Height", "BaseHeightVelocity", "Landing"],
    "PPO_BASIC": ["JointPosition", "JointVelocity", "Pitch", "Height", "BaseHeightVelocity", "Landing"],
    "PPO_BASIC_X": ["JointPosition", "JointVelocity", "Pitch", "Height", "BaseHeightVelocity", "VelocityX", "Landing"],
    "PPO_BASIC_CONTACT": ["JointPosition", "JointVelocity", "Pitch", "Height", "BaseHeightVelocity", "Landing",
                          "BooleanContact"],
    "ARS_BACKFLIP": ["JointPosition", "JointVelocity", "Height", "BaseHeightVelocity", "PitchBackFlip"],
    "PPO_BACKFLIP": ["JointPosition", "JointVelocity", "Height", "BaseHeightVelocity", "PitchBackFlip", "Landing"],
    "PPO_CONTINUOUS_JUMPING_FORWARD": ["JointPosition", "JointVelocity", "Height", "BaseHeightVelocity", "Pitch",
                                       "Landing", "Jumping"],
}


def observation_layout(mode: str, cfg: SimpleNamespace):
    """[(sensor name as in the reference's obs dict, start, stop)], high, low for an obs mode."""
    table = _sensor_table(cfg)
    layout, highs, lows, n = [], [], [], 0
    for key in SENSOR_SETS[mode]:
        name, hi, lo = table[key]
        hi = np.asarray(hi, dtype=np.float64).reshape(-1)
        lo = np.asarray(lo, dtype=np.float64).reshape(-1)
        layout.append((name, n, n + hi.size))
        highs.append(hi)
        lows.append(lo)
        n += hi.size
    return layout, np.concatenate(highs), np.concatenate(lows)
