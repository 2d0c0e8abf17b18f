"""The C-ABI library loads, exports every symbol include/thylacine_b200.h declares, and reports the
reference's assertion failures as status codes.  No compute calls (no GPU needed)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "thylacine_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(thy_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from thylacine_b200 import _capi
    handle = C.CDLL(str(built_lib))
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    # and the ctypes table binds exactly the declared set
    import thylacine_b200.samplers  # noqa: F401  (registers the sampler entry points)
    assert sorted(_capi.SIGNATURES) == declared


def test_version_and_launch_counter(built_lib):
    import thylacine_b200 as t
    assert b"sm_100a" in t.lib().thy_version()
    assert t.kernelLaunchCount() >= 0


def test_reference_assertions_become_status_codes(built_lib):
    import thylacine_b200 as t
    from thylacine_b200 import _capi
    with pytest.raises(t.ThylacineError) as e:       # RecordedData.scala:57
        t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("foo", [1, 2], [3, -5])])
    assert e.value.status == _capi.THY_ERR_INVALID_ARGUMENT
    with pytest.raises(t.ThylacineError) as e:       # UniformDistribution.scala:36-39
        t.UnnormalisedPosterior([t.UniformPrior.fromBounds("u", [1, 1], [0, 1])])
    assert e.value.status == _capi.THY_ERR_INVALID_ARGUMENT
    with pytest.raises(t.ThylacineError) as e:       # UnnormalisedPosterior.scala:35-47 (duplicate labels)
        t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("foo", [1], [1]),
                                 t.GaussianPrior.fromConfidenceIntervals("foo", [2], [1])])
    assert e.value.status == _capi.THY_ERR_INVALID_ARGUMENT
    with pytest.raises(t.ThylacineError) as e:       # IndexedCollection.scala:25-31 (unknown label)
        t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("foo", [1, 2], [3, 5])],
                                [t.GaussianLinearLikelihood.of([[1, 3], [2, 4]], [7, 10], [.01, .01], "bar")])
    assert e.value.status == _capi.THY_ERR_UNKNOWN_LABEL
    with pytest.raises(t.ThylacineError) as e:       # block width != prior dimension
        t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("foo", [1, 2, 3], [3, 5, 1])],
                                [t.GaussianLinearLikelihood.of([[1, 3], [2, 4]], [7, 10], [.01, .01], "foo")])
    assert e.value.status == _capi.THY_ERR_INVALID_ARGUMENT
    with pytest.raises(AssertionError):              # GaussianLinearLikelihood.scala:43-45
        t.GaussianLinearLikelihood.of([[1, 3], [2, 4]], [7, 10, 11], [.01, .01, .01], "foo")
    with pytest.raises(t.ThylacineError) as e:       # non-SPD covariance
        t.UnnormalisedPosterior([t.GaussianPrior.fromCovarianceMatrix("c", [0, 0], [[1, 2], [2, 1]])])
    assert e.value.status == _capi.THY_ERR_INVALID_ARGUMENT


def test_user_links_compile_without_a_gpu_and_report_compiler_errors(built_lib):
    """thy_model_register_link (NonLinearForwardModel.of(evaluation = closure) as CUDA source): NVRTC needs no device, so
    registration -- and its error reporting, thread-local and handle-local -- is testable here."""
    import thylacine_b200 as t
    from thylacine_b200 import _capi
    L = t.lib()
    h = C.c_void_p()
    _capi.check(L.thy_model_create(C.byref(h)))
    out = C.c_int()
    assert L.thy_model_register_link(h, b"z / (1.0 + fabs(z))", b"1.0 / ((1.0 + fabs(z)) * (1.0 + fabs(z)))", C.byref(out)) == 0
    assert out.value == _capi.LINK_USER_BASE
    assert L.thy_model_register_link(h, b"exp(-z * z)", None, C.byref(out)) == 0 and out.value == _capi.LINK_USER_BASE + 1
    assert L.thy_model_register_link(h, b"no_such_function(z)", None, C.byref(out)) == _capi.THY_ERR_INVALID_ARGUMENT
    assert b"no_such_function" in L.thy_last_error()
    buf = C.create_string_buffer(2048)
    assert L.thy_model_last_error(h, buf, 2048) == 0 and b"no_such_function" in buf.value   # handle-local copy
    assert L.thy_model_register_link(h, b"", None, C.byref(out)) == _capi.THY_ERR_INVALID_ARGUMENT
    L.thy_model_destroy(h)


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import thylacine_b200 as t
    from thylacine_b200 import _capi
    with pytest.raises(t.ThylacineError) as e:
        t.UnnormalisedPosterior([t.GaussianPrior.fromConfidenceIntervals("foo", [1, 2], [3, 5])])
    assert e.value.status == _capi.THY_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under thylacine_b200/ may reference it."""
    for path in (ROOT / "thylacine_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"}:
            text = path.read_text()
            assert "ref_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path
