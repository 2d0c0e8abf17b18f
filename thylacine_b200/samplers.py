"""Samplers and integrators built on the batched posterior (HMC, Leapfrog-MCMC, SLQ) -- bindings are
registered here as the C-ABI entry points land."""
from __future__ import annotations
