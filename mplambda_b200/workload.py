"""Synthetic, seeded planner workloads.

Pose i of a stream is a pure function of (seed, i): Philox4x32-10 keyed by the seed, counter
(i, call, 0, 0) -> uniforms u0..u5, mapped through the reference's own sampling rules
(`include/mpl/randomize.hpp:12-38`: Shoemake quaternion, per-axis uniform translation).
Host-side numpy only; this is input generation, not part of the timed path.
"""
import math
import numpy as np

SEED = 0x6D706C62  # "mplb"

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counter_lo, counter_hi, key0, key1=0):
    """Vectorised Philox4x32-10.  counter = (counter_lo, counter_hi, 0, 0) -> 4 uint32 arrays."""
    c0 = np.asarray(counter_lo, dtype=np.uint64) & _MASK
    c1 = np.broadcast_to(np.uint64(counter_hi), c0.shape).astype(np.uint64) & _MASK
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0, k1 = int(key0) & 0xFFFFFFFF, int(key1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(n, seed=SEED, ncols=6, start=0):
    """[n, ncols] float64 uniforms in [0,1), column j of row i = f(seed, start+i, j)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    cols = []
    call = 0
    while len(cols) < ncols:
        cols.extend(philox4x32(idx, call, seed))
        call += 1
    u = np.stack(cols[:ncols], axis=1).astype(np.float64) * (1.0 / 4294967296.0)
    return u


def se3_poses(n, bounds_min, bounds_max, seed=SEED, start=0):
    """[n,7] float64 states (qx,qy,qz,qw,tx,ty,tz) distributed as SE3RigidBodyScenario::
    randomSample (`se3_rigid_body_scenario.hpp:107-113`, `randomize.hpp:12-38`)."""
    u = uniforms(n, seed, 6, start)
    a, b, c = u[:, 0], 2 * math.pi * u[:, 1], 2 * math.pi * u[:, 2]
    lo, hi = np.asarray(bounds_min, dtype=np.float64), np.asarray(bounds_max, dtype=np.float64)
    out = np.empty((n, 7), dtype=np.float64)
    out[:, 0] = np.sqrt(1 - a) * np.sin(b)
    out[:, 1] = np.sqrt(1 - a) * np.cos(b)
    out[:, 2] = np.sqrt(a) * np.sin(c)
    out[:, 3] = np.sqrt(a) * np.cos(c)
    out[:, 4:7] = lo + u[:, 3:6] * (hi - lo)
    return out


def _f32(x):
    """CLI doubles pass through a float in the reference (`src/mpl/demo/app_options.cpp:17-27`)."""
    return [float(np.float32(v)) for v in x]


# Canonical problem parameters of the reference's scripts (scripts/*.sh).
# start/goal are (qx,qy,qz,qw,x,y,z); every number is float-truncated like the reference's CLI.
SE3_SCENES = {
    "twistycool": dict(env="se3/Twistycool_env.dae", robot="se3/Twistycool_robot.dae",
                       start=_f32([0, 1, 0, 0, 270, 160, -200]), goal=_f32([0, 1, 0, 0, 270, 160, -400]),
                       min=_f32([53.46, -21.25, -476.86]), max=_f32([402.96, 269.25, -91.0])),
    "alpha15": dict(env="se3/alpha_env-1.5.dae", robot="se3/alpha_robot.dae",
                    start=_f32([0, 1, 0, 0, -21.91, -4.11, -14.14]), goal=_f32([0, 1, 0, 0, -21.91, -4.11, 68.86]),
                    min=_f32([-281.64, -119.64, -176.86]), max=_f32([189.05, 189.18, 174.86])),
    "apartment": dict(env="se3/Apartment_env.dae", robot="se3/Apartment_robot.dae",
                      start=_f32([0, 0, 0, -1, 241.81, 106.15, 36.46]), goal=_f32([0, 0, 0, -1, -31.19, -99.85, 36.46]),
                      min=_f32([-73.76, -179.59, -0.03]), max=_f32([295.77, 168.26, 90.39])),
    "cubicles": dict(env="se3/cubicles_env.dae", robot="se3/cubicles_robot.dae",
                     start=_f32([0, 1, 0, 0, -4.96, -40.62, 70.57]), goal=_f32([0, 1, 0, 0, 200.0, -40.62, 70.57]),
                     min=_f32([-508.88, -230.13, -123.75]), max=_f32([319.62, 531.87, 101.0])),
    "home": dict(env="se3/Home_env.dae", robot="se3/Home_robot.dae",
                 start=_f32([0, 1, 0, 0, 252.95, -214.95, 46.19]), goal=_f32([0, 1, 0, 0, 262.95, 75.05, 46.19]),
                 min=_f32([-383.802642822, -371.469055176, -0.196851730347]),
                 max=_f32([324.997131348, 337.893371582, 142.332290649])),
}


# ---- Fetch (fetch_robot.hpp:109-175, scripts/fetch_task_00{1,2}.sh) -------------------------------
FETCH_JOINT_MIN = np.array([0.0, -1.6056, -1.221, -np.inf, -2.251, -np.inf, -2.16, -np.inf])
FETCH_JOINT_MAX = np.array([0.38615, 1.6056, 1.518, np.inf, 2.251, np.inf, 2.16, np.inf])
FETCH_SCALE = np.array([10.0, 4, 2, 1, 1, 1, 1, 1])          # FetchScenario::scale (fetch_scenario.hpp:42-45)


def frame_from_option(v):
    """--env-frame x,y,z,rx,ry,rz -> 3x4 [R|t]: axis-angle vector as parsed by
    OptionParser<Eigen::Transform> (app_options.hpp:83-94), Eigen AngleAxis::toRotationMatrix."""
    v = np.array(_f32(v))
    f = np.zeros((3, 4))
    f[:, :3] = np.eye(3)
    axis = v[3:6]
    angle = float(np.sqrt(axis @ axis))
    if abs(angle) > 1e-6:
        a = axis / angle
        s, c = math.sin(angle), math.cos(angle)
        sa, c1 = s * a, (1 - c) * a
        R = np.empty((3, 3))
        t = c1[0] * a[1]; R[0, 1] = t - sa[2]; R[1, 0] = t + sa[2]
        t = c1[0] * a[2]; R[0, 2] = t + sa[1]; R[2, 0] = t - sa[1]
        t = c1[1] * a[2]; R[1, 2] = t - sa[0]; R[2, 1] = t + sa[0]
        for i in range(3):
            R[i, i] = c1[i] * a[i] + c
        f[:, :3] = R
    f[:, 3] = v[:3]
    return f


FETCH_TASKS = {
    "fetch_task_001": dict(env_frame=[0.57, -0.90, 0.00, 0, 0, -1.570796326794897],
                           start=_f32([0.1, 1.570796326794897, 1.570796326794897, 0, 1.570796326794897, 0,
                                       1.570796326794897, 0])),
    "fetch_task_002": dict(env_frame=[0.48, 1.09, 0.00, 0, 0, -1.570796326794897],
                           start=_f32([0.18132, 0.249076, 0.153913, 1.47009, 1.48902, -0.125871, -1.74715, -1.45724])),
}


def fetch_configs(n, seed=SEED, start=0):
    """[n,8] joint vectors distributed as FetchRobot::randomConfig (fetch_robot.hpp:163-175): each joint
    uniform in [max(jointMin,-4pi), min(jointMax,4pi)]."""
    lo = np.maximum(FETCH_JOINT_MIN, -4 * math.pi)
    hi = np.minimum(FETCH_JOINT_MAX, 4 * math.pi)
    return lo + uniforms(n, seed, 8, start) * (hi - lo)
