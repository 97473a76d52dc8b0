"""zisafvm_b200: B200-native (sm_100a) residual path for ZisaFVM-style Euler + gravity solvers.

The package is a thin host mirror of the reference's operator interfaces over ``libzfvm_b200.so``
(hand-written CUDA kernels behind the C ABI of ``include/zfvm.h``).  There is no CPU fallback.
"""
from .grid import (Grid, HybridWENOParams, QRDegrees, StencilFamilies, StencilFamilyParams, WENO_PARAMS,
                   compute_stencil_families, cube_mesh, square_mesh)
from ._capi import ZfvmError
from .solver import (AllVariables, CudaContext, CudaEulerRateOfChange, CudaRungeKutta, EulerParams, FrozenBC, Gravity,
                     LocalCFL)

__all__ = [
    "Grid", "HybridWENOParams", "QRDegrees", "StencilFamilies", "StencilFamilyParams", "WENO_PARAMS",
    "compute_stencil_families", "cube_mesh", "square_mesh", "AllVariables", "CudaContext", "CudaEulerRateOfChange",
    "CudaRungeKutta", "EulerParams", "FrozenBC", "Gravity", "LocalCFL", "ZfvmError",
]
