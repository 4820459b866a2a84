"""numpy float32 model of the device's error-bounded triangle-triangle filter
(mplambda_b200/csrc/se3_kernels.cu: triTriFilter) next to the FP64 specification, used by
tests/test_filter_bound.py to check that the per-axis error bound really covers the FP32-vs-FP64
difference of every separating-axis gap.  Test infrastructure only."""
import numpy as np

F = np.float32
EPS = F(np.finfo(np.float32).eps)


def quat_to_R(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz); R[..., 0, 1] = txy - twz; R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz; R[..., 1, 1] = 1 - (txx + tzz); R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy; R[..., 2, 1] = tyz + twx; R[..., 2, 2] = 1 - (txx + tyy)
    return R


def rel_transform(states):
    """-> (R, T) float64 with R = R1^T, T = -R1^T t1 (env frame -> robot frame)."""
    R1 = quat_to_R(states[..., :4])
    R = np.swapaxes(R1, -1, -2)
    T = -np.einsum("...ij,...j->...i", R, states[..., 4:7])
    return R, T


def cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def axes_and_gaps(p2, p3, q1, q2, q3, e, f):
    """17 axes in FCL order; gap > 0 <=> the axis separates."""
    n1, m1 = cross(e[0], e[1]), cross(f[0], f[1])
    axes = [n1, m1] + [cross(e[i], f[j]) for i in range(3) for j in range(3)] + \
           [cross(e[i], n1) for i in range(3)] + [cross(f[i], m1) for i in range(3)]
    gaps = []
    zero = np.zeros_like(p2[..., 0])
    for ax in axes:
        P = [zero, dot(ax, p2), dot(ax, p3)]
        Q = [dot(ax, q1), dot(ax, q2), dot(ax, q3)]
        mx1, mn1 = np.maximum.reduce(P), np.minimum.reduce(P)
        mx2, mn2 = np.maximum.reduce(Q), np.minimum.reduce(Q)
        gaps.append(np.maximum(mn1 - mx2, mn2 - mx1))
    return axes, np.stack(gaps, axis=-1)


def spec_gaps(P, Q, states):
    """FP64 specification (oracle order: Q' = R Q + T, everything relative to P1)."""
    R, T = rel_transform(states)
    Qp = np.einsum("...ij,...kj->...ki", R, Q) + T[..., None, :]
    p2, p3 = P[..., 1, :] - P[..., 0, :], P[..., 2, :] - P[..., 0, :]
    q = Qp - P[..., 0:1, :]
    q1, q2, q3 = q[..., 0, :], q[..., 1, :], q[..., 2, :]
    e = [p2, p3 - p2, -p3]
    f = [q2 - q1, q3 - q2, q1 - q3]
    return axes_and_gaps(p2, p3, q1, q2, q3, e, f)[1]


def filter_gaps(P, Q, states, env_radius, robot_radius):
    """float32 model of the device filter -> (gaps32 [n,17], E [n,17])."""
    R, T = rel_transform(states)
    R32, T32 = R.astype(F), T.astype(F)
    P1 = P[..., 0, :].astype(F)
    p2 = (P[..., 1, :] - P[..., 0, :]).astype(F)
    p3 = (P[..., 2, :] - P[..., 0, :]).astype(F)
    Q1 = Q[..., 0, :].astype(F)
    F1 = (Q[..., 1, :] - Q[..., 0, :]).astype(F)
    F2 = (Q[..., 2, :] - Q[..., 0, :]).astype(F)

    def rot(v):
        return (R32[..., :, 0] * v[..., 0:1] + R32[..., :, 1] * v[..., 1:2] + R32[..., :, 2] * v[..., 2:3]).astype(F)
    q1 = (rot(Q1) + T32 - P1).astype(F)
    f1, f2 = rot(F1), rot(F2)
    q2, q3 = (q1 + f1).astype(F), (q1 + f2).astype(F)
    e = [p2, (p3 - p2).astype(F), -p3]
    f = [f1, (f2 - f1).astype(F), -f2]
    axes, gaps = axes_and_gaps(p2, p3, q1, q2, q3, e, f)

    qn = (states[..., :4] ** 2).sum(-1)
    scale = np.maximum(1.0, qn)
    tn = np.sqrt((states[..., 4:7] ** 2).sum(-1))
    eta = (8.0 * EPS * (scale * (env_radius + tn) + robot_radius)).astype(F)
    amax = lambda *vs: np.maximum.reduce([np.abs(v).max(-1) for v in vs])
    Me, Mf = amax(*e), amax(*f)
    U = amax(p2, p3, q1, q2, q3)
    S = eta + F(79) * EPS * U
    K = {"n": F(120) * EPS * U * Me * Me, "m": F(792) * EPS * U * Mf * Mf, "ef": F(456) * EPS * U * Me * Mf,
         "g": F(384) * EPS * U * Me * Me * Me, "h": F(2400) * EPS * U * Mf * Mf * Mf}
    kinds = ["n", "m"] + ["ef"] * 9 + ["g"] * 3 + ["h"] * 3
    E = np.stack([np.abs(ax).sum(-1) * S + K[k] for ax, k in zip(axes, kinds)], axis=-1)
    return gaps, E.astype(F)


# ---- Fetch: box-triangle separating-axis cull (mplambda_b200/csrc/fetch_kernels.cu: boxTriangleApart) ----------------
CULL_SLACK = np.float32(1e-4)


def box_triangle_apart32(e, q):
    """float32 model of boxTriangleApart: e [n,3] box half extents, q [n,3,3] triangle vertices in the box's frame.
    True = the device does not list the leaf for the exact test."""
    f32 = np.float32
    e = e.astype(f32); q = q.astype(f32)
    mn, mx = q.min(1), q.max(1)
    apart = ((mn > e + CULL_SLACK) | (mx < -e - CULL_SLACK)).any(1)
    ed = np.stack([q[:, 1] - q[:, 0], q[:, 2] - q[:, 1], q[:, 0] - q[:, 2]], 1)
    n = np.cross(ed[:, 0], ed[:, 1]).astype(f32)
    d = (n * q[:, 0]).sum(1, dtype=f32)
    a = np.abs(n)
    apart |= np.abs(d) > (e * a).sum(1, dtype=f32) + CULL_SLACK * a.sum(1, dtype=f32)
    for j in range(3):
        fx, fy, fz = ed[:, j, 0], ed[:, j, 1], ed[:, j, 2]
        z = np.zeros_like(fx)
        for L in (np.stack([z, -fz, fy], 1), np.stack([fz, z, -fx], 1), np.stack([-fy, fx, z], 1)):
            p = np.einsum("nk,nvk->nv", L, q).astype(f32)
            al = np.abs(L)
            r = (e * al).sum(1, dtype=f32) + CULL_SLACK * al.sum(1, dtype=f32)
            apart |= (p.min(1) > r) | (p.max(1) < -r)
    return apart


def box_triangle_gap64(e, q):
    """float64 specification: the largest separating gap in METRES over the 13 axes (normalised axes; > 0 iff the box
    and the triangle are disjoint -- the 13 axes are necessary and sufficient for a box and a triangle)."""
    e = e.astype(np.float64); q = q.astype(np.float64)
    gaps = []
    mn, mx = q.min(1), q.max(1)
    gaps.append(np.maximum(mn - e, -mx - e).max(1))
    ed = np.stack([q[:, 1] - q[:, 0], q[:, 2] - q[:, 1], q[:, 0] - q[:, 2]], 1)
    axes = [np.cross(ed[:, 0], ed[:, 1])]
    eye = np.eye(3)
    for j in range(3):
        for i in range(3):
            axes.append(np.cross(np.broadcast_to(eye[i], ed[:, j].shape), ed[:, j]))
    for L in axes:
        ln = np.linalg.norm(L, axis=1)
        ok = ln > 1e-12
        Ln = L / np.where(ok, ln, 1.0)[:, None]
        p = np.einsum("nk,nvk->nv", Ln, q)
        r = (e * np.abs(Ln)).sum(1)
        g = np.maximum(p.min(1) - r, -p.max(1) - r)
        gaps.append(np.where(ok, g, -np.inf))
    return np.max(np.stack(gaps, 1), 1)
