"""mogp_emulator_b200 -- B200-native (sm_100a) GP fit+predict hot path of mogp-emulator behind the reference's
``GaussianProcessGPU`` / ``MultiOutputGP_GPU`` Python API (ctypes over libmogp_b200.so; no CPU fallback)."""
from .libmogp import gpu_usable, HAVE_LIBMOGP
from .kernels import SquaredExponential, Matern52
from .hyper import GPParams, GPPriors, InvGammaPrior, GammaPrior, LogNormalPrior, WeakPrior
from .GaussianProcessGPU import GaussianProcessGPU, PredictResult, GPUUnavailableError
from .MultiOutputGP_GPU import MultiOutputGP_GPU
from .fitting import fit_GP_MAP
from . import validation
from .HistoryMatching import HistoryMatching
from .SequentialDesign import MICEFastGP

__all__ = ["gpu_usable", "HAVE_LIBMOGP", "SquaredExponential", "Matern52", "GPParams", "GPPriors", "InvGammaPrior",
           "GammaPrior", "LogNormalPrior", "WeakPrior", "GaussianProcessGPU", "MultiOutputGP_GPU", "PredictResult",
           "GPUUnavailableError", "fit_GP_MAP", "validation", "HistoryMatching", "MICEFastGP"]
