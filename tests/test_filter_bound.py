"""CPU: the device's FP32 triangle-triangle filter is only allowed to answer when its FP32 gap
exceeds the per-axis error bound E; this checks E against the FP64 specification on random and
near-touching configurations with mixed triangle sizes and large translations (numpy float32
model in tests/filter_model.py mirrors mplambda_b200/csrc/se3_kernels.cu: triTriFilter)."""
import numpy as np
import pytest

from filter_model import filter_gaps, rel_transform, spec_gaps


def cases(rng, n, robot_size, env_size, offset, near):
    P = rng.normal(size=(n, 3, 3)) * robot_size + rng.normal(size=(n, 1, 3)) * robot_size * 2
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    st = np.concatenate([q, rng.normal(size=(n, 3)) * offset], 1)
    R, T = rel_transform(st)
    Qr = rng.normal(size=(n, 3, 3)) * env_size + P.mean(1, keepdims=True) + rng.normal(size=(n, 1, 3)) * near
    Q = np.einsum("nji,nkj->nki", R, Qr - T[:, None, :])   # back to the env frame
    return P, Q, st


@pytest.mark.parametrize("cfg", [(10, 10, 500, 1), (10, 300, 500, 5), (30, 400, 1000, 0.1), (1, 1, 10, 0.01),
                                 (50, 50, 0, 1), (5, 500, 2000, 0.5), (0.01, 300, 300, 0.01), (100, 0.1, 1e4, 1)])
def test_error_bound_covers_fp32_error(cfg):
    rng = np.random.default_rng(hash(cfg) % 2**32)
    P, Q, st = cases(rng, 50000, *cfg)
    env_r, rob_r = np.linalg.norm(Q, axis=-1).max(), np.linalg.norm(P, axis=-1).max()
    g64 = spec_gaps(P, Q, st)
    g32, E = filter_gaps(P, Q, st, env_r, rob_r)
    err = np.abs(g32.astype(np.float64) - g64)
    assert (err <= 0.5 * E).all(), "bound violated: max err/E = %g" % (err / E).max()
    # the filter's certain answers agree with the specification
    sep = (g32 > E).any(1)
    unsure = ~sep & (g32 >= -E).any(1)
    spec_hit = ~(g64 > 0).any(1)
    assert (spec_hit[sep] == False).all() and (spec_hit[~sep & ~unsure] == True).all()  # noqa: E712
    if cfg[2] <= 2000:      # with a sane translation magnitude the exact path is rarely needed
        assert unsure.mean() < 0.02


def test_degenerate_parallel_edges_are_unsure_not_wrong():
    # axis-aligned boxes at identity: many e x f axes vanish; the filter must not decide on them
    P = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0.0]]])
    Q = np.array([[[2, 0, 0], [3, 0, 0], [2, 1, 0.0]]])
    st = np.array([[0, 0, 0, 1, 0, 0, 0.0]])
    g64 = spec_gaps(P, Q, st)
    g32, E = filter_gaps(P, Q, st, 4.0, 2.0)
    assert (np.abs(g32 - g64) <= E).all()
    assert (g32 > E).any()       # separated along x with margin 1: certain


def _box_triangle_cases(rng, n, size, spread, near):
    from filter_model import box_triangle_apart32, box_triangle_gap64  # noqa: F401
    e = rng.uniform(0.02, 0.35, size=(n, 3))                       # the Fetch shapes' boxes: 2 .. 35 cm half extents
    c = rng.normal(size=(n, 1, 3)) * spread                        # triangle somewhere around the box ...
    q = c + rng.normal(size=(n, 3, 3)) * size
    # ... and a share of them pulled to the box's surface, where the decision is close
    k = n // 2
    s = np.sign(rng.normal(size=(k, 3)))
    q[:k] += (s * e[:k] * (1 + rng.normal(size=(k, 1)) * near))[:, None, :] - c[:k]
    return e, q


@pytest.mark.parametrize("cfg", [(0.05, 0.5, 1e-3), (0.5, 1.0, 1e-4), (2.0, 0.5, 1e-2), (0.01, 0.3, 1e-5), (5.0, 3.0, 1e-3)])
def test_box_triangle_cull_is_conservative(cfg):
    """fetch_kernels.cu boxTriangleApart (FP32, slack 1e-4 m) may only drop a leaf when the FP64 13-axis test says box
    and triangle are disjoint -- by the slack less the FP32 error, which must stay far below it at robot scale."""
    from filter_model import box_triangle_apart32, box_triangle_gap64
    rng = np.random.default_rng(hash(cfg) % 2**32)
    e, q = _box_triangle_cases(rng, 200000, *cfg)
    apart = box_triangle_apart32(e, q)
    gap = box_triangle_gap64(e, q)
    assert apart.any() and (~apart).any()
    # never drops an intersecting or touching pair, and everything it drops is clear by most of the slack
    assert (gap[apart] > 0).all()
    assert gap[apart].min() > 0.5e-4, gap[apart].min()
    # and it is not vacuous: pairs that are clear by a millimetre in the L1-scaled sense are dropped
    assert apart[gap > 2e-3 * np.sqrt(3)].mean() > 0.99


def test_box_triangle_cull_degenerate_triangles():
    from filter_model import box_triangle_apart32
    e = np.array([[0.1, 0.1, 0.1]] * 3)
    q = np.array([[[0.05, 0, 0]] * 3,                                   # a point inside the box
                  [[0.3, 0, 0], [0.3, 0, 0], [0.4, 0, 0]],              # a segment outside
                  [[-1, 0, 0], [1, 0, 0], [0, 0, 0.0]]])                # a collinear triangle through the box
    assert box_triangle_apart32(e, q).tolist() == [False, True, False]
