"""gpyopt_b200: B200-native engine for GPyOpt's GP-posterior + acquisition hot path.

Host side (this package): the reference's plugin interfaces re-hosted on a ctypes binding of
``lib/libgpo_b200.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/gpo_b200.h``).
There is no CPU fallback: importing the compute classes without the built library, or constructing an
engine without a B200, raises.
"""
from ._lib import Engine, GPOError, load_library, LIB_PATH  # noqa: F401
from . import kern  # noqa: F401
from .gp import GPRegressionB200, GammaPrior, HMC  # noqa: F401
from .models import GPModel, GPModel_MCMC  # noqa: F401
from .acquisitions import (AcquisitionEI, AcquisitionMPI, AcquisitionLCB, AcquisitionEI_MCMC,  # noqa: F401
                           AcquisitionMPI_MCMC, AcquisitionLCB_MCMC, AcquisitionLP)

__version__ = '0.1.0'
