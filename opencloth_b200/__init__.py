"""opencloth_b200 — B200-native (sm_100a) Verlet mass-spring cloth step of mmmovania/opencloth.

Only the hot path: StepPhysics = ComputeForces -> IntegrateVerlet -> EllipsoidCollision
(/root/reference/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp:557-562), behind the C-ABI of
include/opencloth.h.  Importing the package does not need a GPU; creating a ``Cloth`` does, and
fails loudly without one (there is no CPU fallback).
"""
from ._abi import (LIB_PATH, OcParams, OpenClothError, OC_KERNEL_AUTO, OC_KERNEL_GATHER, OC_KERNEL_MARCH, OC_KERNEL_MARCH2, OC_KERNEL_RESIDENT, OC_KERNEL_TWIN, OC_KERNEL_STREAM, OC_KERNEL_STREAM2, OC_KERNEL_BANDRES,
                   OC_INTEGRATOR_VERLET, OC_INTEGRATOR_EULER, OC_INTEGRATOR_SEMI_IMPLICIT)
from .cloth import Cloth, default_params, version, link_bands_local

__all__ = ["Cloth", "default_params", "version", "link_bands_local", "OcParams", "OpenClothError", "LIB_PATH",
           "OC_KERNEL_AUTO", "OC_KERNEL_GATHER", "OC_KERNEL_MARCH", "OC_KERNEL_MARCH2", "OC_KERNEL_RESIDENT", "OC_KERNEL_TWIN", "OC_KERNEL_STREAM", "OC_KERNEL_STREAM2", "OC_KERNEL_BANDRES"]
