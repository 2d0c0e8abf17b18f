"""ctypes binding of ``include/thylacine_b200.h`` (the C ABI of libthylacine_b200.so).

There is no CPU fallback: if the shared library is missing this module raises at import of the
symbol table, and every evaluation raises ``ThylacineError`` (THY_ERR_NO_DEVICE) without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libthylacine_b200.so"

THY_OK = 0
STATUS_NAMES = {
    0: "THY_OK", 1: "THY_ERR_INVALID_ARGUMENT", 2: "THY_ERR_UNKNOWN_LABEL", 3: "THY_ERR_CUDA", 4: "THY_ERR_NCCL",
    5: "THY_ERR_STATE", 6: "THY_ERR_UNSUPPORTED", 7: "THY_ERR_NO_DEVICE",
}
THY_ERR_INVALID_ARGUMENT, THY_ERR_UNKNOWN_LABEL, THY_ERR_CUDA, THY_ERR_NCCL = 1, 2, 3, 4
THY_ERR_STATE, THY_ERR_UNSUPPORTED, THY_ERR_NO_DEVICE = 5, 6, 7

DIST_GAUSSIAN, DIST_CAUCHY, DIST_UNIFORM = 0, 1, 2
LINK_USER_BASE = 100
LINK_IDENTITY, LINK_TANH, LINK_EXP, LINK_SQUARE, LINK_SOFTPLUS, LINK_SIGMOID = range(6)
JAC_CONSTANT, JAC_ANALYTIC, JAC_FINITE_DIFFERENCE = 0, 1, 2
OPT_CAUCHY_REFERENCE_SIGN, OPT_FD_INCREMENTAL, OPT_KERNEL_PATH, OPT_FUSE_PRIOR = 0, 1, 2, 3
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED_DMMA, KERNEL_STREAM, KERNEL_WIDE, KERNEL_FUSED_STREAMING = 0, 1, 2, 3, 4, 5


class ThylacineError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class LikelihoodDesc(C.Structure):
    _fields_ = [
        ("distribution", C.c_int), ("link", C.c_int), ("jacobian", C.c_int), ("n", C.c_int), ("nblocks", C.c_int),
        ("labels", C.POINTER(C.c_char_p)), ("dims", C.POINTER(C.c_int)),
        ("blocks", C.POINTER(C.POINTER(C.c_double))), ("offset", C.POINTER(C.c_double)),
        ("measurements", C.POINTER(C.c_double)), ("uncertainties", C.POINTER(C.c_double)),
        ("differential", C.c_double),
    ]


class HmcmcConfigC(C.Structure):
    _fields_ = [("stepsBetweenSamples", C.c_int), ("stepsInDynamicsSimulation", C.c_int),
                ("warmupStepCount", C.c_int), ("dynamicsSimulationStepSize", C.c_double)]


class HmcStatsC(C.Structure):
    _fields_ = [("jumpAttempts", C.c_int64), ("jumpAcceptances", C.c_int64), ("lastHamiltonianDifferential", C.c_double)]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes).  tests/test_abi.py checks this table against the header.
SIGNATURES = {
    "thy_version": (C.c_char_p, []),
    "thy_last_error": (C.c_char_p, []),
    "thy_kernel_launch_count": (C.c_uint64, []),
    "thy_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "thy_set_device": (C.c_int, [C.c_int]),
    "thy_model_last_error": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "thy_model_create": (C.c_int, [C.POINTER(_vp)]),
    "thy_model_destroy": (C.c_int, [_vp]),
    "thy_model_add_gaussian_prior_ci": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp, _dp]),
    "thy_model_add_gaussian_prior_cov": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp, _dp]),
    "thy_model_add_uniform_prior": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp, _dp]),
    "thy_model_add_cauchy_prior": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp, _dp]),
    "thy_model_add_likelihood": (C.c_int, [_vp, C.POINTER(LikelihoodDesc)]),
    "thy_model_register_link": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]),
    "thy_model_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "thy_model_finalize": (C.c_int, [_vp]),
    "thy_model_dimension": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "thy_model_label_layout": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "thy_model_set_stream": (C.c_int, [_vp, _vp]),
    "thy_model_synchronize": (C.c_int, [_vp]),
    "thy_logpdf_batch": (C.c_int, [_vp, C.c_int64, _dp, _dp]),
    "thy_logpdf_grad_batch": (C.c_int, [_vp, C.c_int64, _dp, _dp, _dp]),
    "thy_logpdf_grad_batch_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp]),
    "thy_logpdf_grad_batch_async": (C.c_int, [_vp, C.c_int64, _dp, _dp, _dp, C.POINTER(C.c_int64)]),
    "thy_model_wait": (C.c_int, [_vp, C.c_int64]),
    "thy_model_last_kernel_path": (C.c_char_p, [_vp]),
    "thy_model_enable_timing": (C.c_int, [_vp, C.c_int]),
    "thy_model_kernel_time": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "thy_model_collapse": (C.c_int, [_vp, C.POINTER(_vp)]),
    "thy_model_gaussian_posterior": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "thy_sample_priors": (C.c_int, [_vp, C.c_int64, C.c_uint64, _dp]),
    "thy_sample_priors_dev": (C.c_int, [_vp, C.c_int64, C.c_uint64, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libthylacine_b200.so (built by ``__graft_entry__.build()`` / ``make -C thylacine_b200/csrc``)."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("THYLACINE_B200_LIB", _LIB_PATH))
        if not path.exists():
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(thylacine_b200 has no CPU fallback)")
        handle = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def register(name: str, restype, argtypes) -> None:
    """Used by the sampler modules to extend the symbol table before the library is loaded."""
    SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(status: int) -> None:
    if status != THY_OK:
        raise ThylacineError(status, lib().thy_last_error().decode())
