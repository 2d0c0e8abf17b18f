"""The oracle's restatement of the reference's cubical cover (oracle/slq_cubical_cover.py) -- CPU only.

The reference holds no test for integration/slq/* (SURVEY.md 8c), so the restatement is checked against what the
reference's code guarantees by construction and against hand-worked cases of PointInInterval.findDisjointBoundary
(PointInInterval.scala:82-109)."""
import numpy as np
import pytest

from oracle import slq_cubical_cover as cc


def test_find_disjoint_boundary_cases():
    # overlapping intervals around 0 and 1: the boundary is the midpoint (:91-92)
    assert cc._find_disjoint_boundary(0.0, -2.0, 2.0, 1.0, -1.0, 3.0) == ((-2.0, 0.5), (0.5, 3.0))
    # the smaller interval ends before the midpoint: boundary = its upper bound (:93-94)
    assert cc._find_disjoint_boundary(0.0, -1.0, 0.25, 1.0, -1.0, 3.0) == ((-1.0, 0.25), (0.25, 3.0))
    # the larger interval starts after the midpoint: boundary = its lower bound (:95-96)
    assert cc._find_disjoint_boundary(1.0, 0.75, 3.0, 0.0, -2.0, 2.0) == ((0.75, 3.0), (-2.0, 0.75))
    with pytest.raises(RuntimeError):
        cc._find_disjoint_boundary(0.0, -1.0, 0.2, 1.0, 0.8, 2.0)


@pytest.mark.parametrize("P,d,seed", [(2, 1, 0), (12, 2, 1), (40, 3, 2), (25, 6, 3)])
def test_cover_is_disjoint_point_centred_and_inside_the_initial_bounds(P, d, seed):
    pts = np.random.default_rng(seed).standard_normal((P, d))
    lo0, up0 = cc.initialise_cover(pts)
    # initialise (PointInCubeCollection.scala:60-85): half-width = largest distance to another point per dimension
    for i in range(P):
        assert np.allclose(up0[i] - pts[i], np.max(np.abs(pts - pts[i]), axis=0))
    lo, up = cc.ready_for_sampling(pts)
    assert np.all(lo < up) and np.all(lo <= pts) and np.all(pts <= up)
    assert np.allclose(up - pts, pts - lo)                                   # symmetrize (PointInInterval.scala:51-61)
    assert np.all(lo >= lo0 - 1e-15) and np.all(up <= up0 + 1e-15)            # splits only ever shrink a cube
    for i in range(P):
        for j in range(i + 1, P):
            assert not cc._intersecting(lo[i], up[i], lo[j], up[j])          # makeDisjoint (PointInCube.scala:84-115)
    vol = cc.cube_volumes(lo, up)
    assert np.allclose(vol, np.sum(up - lo, axis=1))                         # "volume" = sum of side lengths (:30-33)
    assert cc.choose_cube(vol, 0.0) == 0 and cc.choose_cube(vol, 0.999999999) == P - 1


def test_sampling_loop_retires_in_ascending_order_and_rebuilds():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-3, 3, (30, 2))
    draws = {}

    def uniform(stream, n):
        return draws.setdefault((stream, n), float(np.random.default_rng(1000003 * stream + n).random()))

    ret, pool, lps, nprop = cc.run_reference_slq(pts, lambda x: -0.5 * float(x @ x), uniform, max_iterations=150, batch=4,
                                                 rebuild_streak=25)
    assert ret.size == 150 and np.all(np.diff(ret) >= 0)                     # the pool minimum only ever rises
    assert np.all(lps > ret[-1] - 1e-15) and nprop >= 150
    assert np.allclose(lps, [-0.5 * float(x @ x) for x in pool])
