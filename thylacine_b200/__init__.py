"""thylacine_b200 -- B200-native (sm_100a) posterior-evaluation hot path behind the reference's API surface.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``include/thylacine_b200.h``), ``_capi`` (ctypes binding) and ``api`` (host-side mirror of the
reference's Scala classes).  There is no CPU fallback.
"""
from ._capi import ThylacineError, lib  # noqa: F401
from . import _capi as capi  # noqa: F401
from .api import (  # noqa: F401
    CauchyLikelihood, CauchyPrior, GaussianLikelihood, GaussianLinearLikelihood, GaussianPrior, LinearForwardModel,
    NonLinearForwardModel, UniformLikelihood, UniformPrior, UnnormalisedPosterior, UserLink, kernelLaunchCount,
)
from .samplers import HmcmcConfig, HmcmcSampledPosterior, HmcmcTelemetryUpdate  # noqa: F401,E402
from .samplers import LeapfrogMcmcConfig, LeapfrogMcmcSampledPosterior  # noqa: F401,E402
from .samplers import SampleMoments, SlqConfig, SlqIntegratedPosterior, SlqTelemetryUpdate  # noqa: F401,E402
from .distributed import Communicator, shard_range  # noqa: F401,E402
from .optimisers import (  # noqa: F401,E402
    ConjugateGradientConfig, ConjugateGradientOptimisedPosterior, CoordinateSlideConfig, CoordinateSlideOptimisedPosterior,
    HookeAndJeevesConfig, HookeAndJeevesOptimisedPosterior, MdsConfig, MdsOptimisedPosterior, OptimisationTelemetryUpdate,
)
from .analytic import GaussianAnalyticPosterior  # noqa: F401,E402
from .visualisation import CartesianSurface  # noqa: F401,E402
