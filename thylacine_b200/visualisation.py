"""``CartesianSurface``: density of sample curves on an x-y grid.

Reference: /root/reference/src/main/scala/thylacine/visualisation/CartesianSurface.scala
    :160-199  ``CartesianSurface.of(xAbscissa, yAbscissa, progressSetCallback, progressIncrementCallback,
              progressFinishCallback)``
    :129-150  ``addSamples(abscissa, samples, ds, kernelVariance)``
    :152-157  ``getResults`` -> Map[(x, y) -> value] after column normalisation
The accumulation runs in the C library (csrc/surface.cu); the progress callbacks are fired here, in the order the
reference fires them (set(n), one increment per sample, one more increment, finish -- :139-147).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np

from ._capi import check, lib, register

_vp, _dp = C.c_void_p, C.POINTER(C.c_double)
register("thy_surface_create", C.c_int, [C.c_int, _dp, C.c_int, _dp, C.POINTER(_vp)])
register("thy_surface_destroy", C.c_int, [_vp])
register("thy_surface_add_samples", C.c_int, [_vp, C.c_int, _dp, C.c_int64, _dp, C.c_double, C.c_double])
register("thy_surface_add_samples_dev", C.c_int, [_vp, C.c_int, _dp, C.c_int64, _vp, C.c_double, C.c_double])
register("thy_surface_results", C.c_int, [_vp, C.c_int, _dp])


def _arr(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class CartesianSurface:
    def __init__(self, *a, **k):
        raise TypeError("use CartesianSurface.of(...), as in the reference")

    @classmethod
    def of(cls, xAbscissa: Sequence[float], yAbscissa: Sequence[float],
           progressSetCallback: Optional[Callable[[int], None]] = None,
           progressIncrementCallback: Optional[Callable[[], None]] = None,
           progressFinishCallback: Optional[Callable[[], None]] = None) -> "CartesianSurface":
        self = object.__new__(cls)
        self._lib = lib()
        self.xAbscissa, self.yAbscissa = _arr(xAbscissa).ravel(), _arr(yAbscissa).ravel()
        self._set, self._inc, self._fin = progressSetCallback, progressIncrementCallback, progressFinishCallback
        h = _vp()
        check(self._lib.thy_surface_create(self.xAbscissa.size, _ptr(self.xAbscissa), self.yAbscissa.size,
                                           _ptr(self.yAbscissa), C.byref(h)))
        self._h = h
        return self

    def addSamples(self, abscissa: Sequence[float], samples, ds: float, kernelVariance: float) -> None:
        """``samples``: (n, m) array-like of curve values over ``abscissa`` (m), or a CUDA fp64 torch tensor."""
        ab = _arr(abscissa).ravel()
        on_device = hasattr(samples, "is_cuda") and samples.is_cuda
        n = int(samples.shape[0]) if hasattr(samples, "shape") else len(samples)
        if n == 0:
            return                                                    # case _ => Async[F].unit (:148-149)
        if self._set:
            self._set(n)
        if on_device:
            assert samples.dim() == 2 and samples.shape[1] == ab.size and samples.is_contiguous()
            check(self._lib.thy_surface_add_samples_dev(self._h, ab.size, _ptr(ab), n, samples.data_ptr(), float(ds),
                                                        float(kernelVariance)))
        else:
            s = _arr(samples).reshape(n, -1)
            assert s.shape[1] == ab.size, "every sample needs one value per abscissa point"
            check(self._lib.thy_surface_add_samples(self._h, ab.size, _ptr(ab), n, _ptr(s), float(ds), float(kernelVariance)))
        if self._inc:
            for _ in range(n + 1):
                self._inc()
        if self._fin:
            self._fin()

    def getResultsArray(self, normaliseColumns: bool = True) -> np.ndarray:
        """values[ix, iy] (column = fixed x, normalised over y as ``getResults`` does)."""
        out = np.empty((self.xAbscissa.size, self.yAbscissa.size))
        check(self._lib.thy_surface_results(self._h, 1 if normaliseColumns else 0, _ptr(out)))
        return out

    def getResults(self) -> Dict[Tuple[float, float], float]:
        v = self.getResultsArray(True)
        return {(float(x), float(y)): float(v[i, j]) for i, x in enumerate(self.xAbscissa) for j, y in enumerate(self.yAbscissa)}

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.thy_surface_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
