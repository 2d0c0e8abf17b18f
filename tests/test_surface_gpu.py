"""CartesianSurface on the device (csrc/surface.cu, through the C ABI) against the oracle's literal restatement of
visualisation/CartesianSurface.scala.  fp64, tolerance 1e-12 relative to the largest grid value (the device multiplies
by reciprocals of xScale / yScale / kernelVariance and sums in a different order)."""
import numpy as np
import pytest

import thylacine_b200 as t
from thylacine_b200 import capi
from oracle import ref_oracle as o

pytestmark = pytest.mark.gpu


def _curves(rng, n, m):
    ab = np.sort(rng.uniform(0.0, 10.0, m))
    return ab, np.cumsum(rng.standard_normal((n, m)) * 0.7, axis=1) + 2.0


@pytest.mark.parametrize("nx,ny,n,m,ds", [(7, 5, 3, 4, 0.37), (33, 20, 11, 17, 0.21), (4, 3, 1, 2, 5.0), (16, 16, 40, 9, 1.3)])
def test_surface_matches_the_oracle(nx, ny, n, m, ds):
    rng = np.random.default_rng(nx * 100 + m)
    ab, samples = _curves(rng, n, m)
    xa = np.linspace(-1.0, 11.0, nx)
    ya = rng.permutation(np.linspace(samples.min() - 1, samples.max() + 1, ny))      # unsorted y: the trapezoid sorts
    surf = t.CartesianSurface.of(xa, ya)
    ref = o.CartesianSurface(xa, ya)
    kv = 0.004
    surf.addSamples(ab, samples, ds, kv)
    ref.add_samples(ab, samples, ds, kv)
    raw, want_raw = surf.getResultsArray(False), ref.results_array(False)
    assert np.max(np.abs(raw - want_raw)) <= 1e-12 * max(1.0, np.max(want_raw))
    got, want = surf.getResultsArray(True), ref.results_array(True)
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(want))
    # a second block accumulates on top of the first (scalarValues.modifyF, CartesianSurface.scala:90-96)
    surf.addSamples(ab, samples[:1] + 0.5, ds, kv)
    ref.add_samples(ab, samples[:1] + 0.5, ds, kv)
    assert np.max(np.abs(surf.getResultsArray(False) - ref.results_array(False))) <= 1e-12 * max(1.0, np.max(want_raw))
    d = surf.getResults()
    assert d[(float(xa[2]), float(ya[1]))] == pytest.approx(ref.get_results()[(float(xa[2]), float(ya[1]))], rel=1e-10, abs=1e-300)
    surf.close()


def test_surface_takes_a_sampler_buffer_on_the_device_and_fires_the_callbacks():
    import torch
    rng = np.random.default_rng(4)
    ab, samples = _curves(rng, 300, 25)
    xa, ya = np.linspace(0, 10, 64), np.linspace(samples.min(), samples.max(), 48)
    calls = []
    surf = t.CartesianSurface.of(xa, ya, lambda n: calls.append(("set", n)), lambda: calls.append("inc"), lambda: calls.append("fin"))
    surf.addSamples(ab, torch.from_numpy(samples).cuda(), 0.25, 0.002)
    host = t.CartesianSurface.of(xa, ya)
    host.addSamples(ab, samples, 0.25, 0.002)
    assert np.array_equal(surf.getResultsArray(False), host.getResultsArray(False))  # deterministic, same path
    assert calls[0] == ("set", 300) and calls.count("inc") == 301 and calls[-1] == "fin"
    cols = surf.getResultsArray(True)
    for i in range(xa.size):
        assert o.trapezoidal_quadrature(ya, cols[i]) == pytest.approx(1.0, rel=1e-12)
    surf.close()
    host.close()


def test_surface_argument_checks_follow_the_reference():
    surf = t.CartesianSurface.of([0.0, 1.0], [0.0, 0.0, 1.0])                        # repeated y value
    surf.addSamples([0.0, 1.0], [[0.1, 0.9]], 0.1, 0.01)
    with pytest.raises(t.ThylacineError) as e:                                       # trapezoidalQuadrature throws
        surf.getResultsArray(True)
    assert e.value.status == capi.THY_ERR_INVALID_ARGUMENT
    assert np.all(np.isfinite(surf.getResultsArray(False)))
    with pytest.raises(t.ThylacineError):
        surf.addSamples([0.0, 1.0], [[0.1, 0.9]], 0.0, 0.01)                         # ds = 0 never terminates upstream
    surf.addSamples([0.0, 1.0], np.empty((0, 2)), 0.1, 0.01)                         # no samples: no-op
    surf.addSamples([0.0], [[0.1]], 0.1, 0.01)                                       # fewer than 2 points: empty chain
    surf.close()
