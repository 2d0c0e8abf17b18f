"""CPU checks of the oracle's CartesianSurface restatement (visualisation/CartesianSurface.scala and helpers).
The reference has no test for this class, so the restatement is pinned by hand-computable cases."""
import math

import numpy as np
import pytest

from oracle import ref_oracle as o


def test_interpolation_points_prepend_and_carry_the_remainder():
    # a 3-4-5 segment sampled every 2 units: t = 0, 0.4, 0.8; remainder 6 - 5 = 1 (Simplex.scala:36-47)
    pts, rem = o.simplex_interpolation_points((0.0, 0.0), (3.0, 4.0), 0.0, 2.0)
    assert rem == pytest.approx(1.0)
    assert pts == [(pytest.approx(2.4), pytest.approx(3.2)), (pytest.approx(1.2), pytest.approx(1.6)), (0.0, 0.0)]
    # the next segment starts 1 unit in
    pts2, rem2 = o.simplex_interpolation_points((3.0, 4.0), (3.0, 8.0), rem, 2.0)
    assert [p[1] for p in pts2] == [pytest.approx(7.0), pytest.approx(5.0)] and rem2 == pytest.approx(1.0)
    # zero-length segment: no point, the remainder is unchanged
    assert o.simplex_interpolation_points((1.0, 1.0), (1.0, 1.0), 0.5, 2.0) == ([], 0.5)


def test_chain_points_drop_the_first_point_of_the_last_segment():
    # two horizontal unit segments, ds = 0.5: points x = 0, 0.5 | 1.0, 1.5; vector = [0.5, 0, 1.5, 1.0] -> drop 1.0
    pts = o.simplex_chain_points([0.0, 1.0, 2.0], [3.0, 3.0, 3.0], 0.5)
    assert sorted(p[0] for p in pts) == [0.0, 0.5, 1.5]
    assert o.simplex_chain_points([0.0, 1.0], [0.0, 0.0], 5.0) == []       # one point only: empty chain
    assert o.simplex_chain_points([0.0], [0.0], 0.1) == []


def test_trapezoid_sorts_and_rejects_repeated_abscissa():
    assert o.trapezoidal_quadrature([2.0, 0.0, 1.0], [4.0, 0.0, 1.0]) == pytest.approx(0.5 * 1 + 0.5 * 5)
    with pytest.raises(RuntimeError):
        o.trapezoidal_quadrature([0.0, 0.0, 1.0], [1.0, 1.0, 1.0])
    with pytest.raises(RuntimeError):
        o.trapezoidal_quadrature([0.0], [1.0])


def test_surface_accumulates_gaussian_kernels_and_normalises_columns():
    s = o.CartesianSurface([0.0, 1.0, 2.0], [0.0, 1.0, 2.0, 4.0])
    s.add_samples([0.0, 1.0, 2.0], [[1.0, 1.0, 1.0]], 0.5, 0.05)           # kept points: x = 0, 0.5, 1.5 at y = 1
    want = sum(math.exp(-0.5 * (((px - 1.0) / 2.0) ** 2 + ((1.0 - 2.0) / 4.0) ** 2) / 0.05) for px in (0.0, 0.5, 1.5))
    assert s.values[(1.0, 2.0)] == pytest.approx(want, rel=1e-15)
    res = s.results_array(True)
    for i in range(3):                                                       # every column integrates to one over y
        assert o.trapezoidal_quadrature(s.y, res[i]) == pytest.approx(1.0, rel=1e-13)
    empty = o.CartesianSurface([0.0, 1.0], [0.0, 1.0])
    assert np.all(empty.results_array(True) == 0.0)                          # integration == 0: values left as they are
