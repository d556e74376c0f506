"""B200-native drop-in for the neural-posterior-estimation hot path of FlopPITy (sbi NSF + SNPE-C).

    from floppity_deprecated_b200 import posterior_nn, SNPE_C, utils      # replaces the sbi imports of
    prior = utils.BoxUniform(low, high, device="cuda")                    # /root/reference/modules.py:1-8
    inference = SNPE_C(prior=prior, density_estimator=posterior_nn(model="nsf", ...), device="cuda")

Everything below these names runs in hand-written sm_100a CUDA kernels (``libfloppity_b200.so``); there is no
CPU fallback and the package never imports ``oracle/``.
"""
from . import _lib
from .flow import NSFFlow, build_nsf, posterior_nn
from .inference import BoxUniform, DirectPosterior, SNPE_C
from .density_estimator import NSFDensityEstimator
from .embedding import CNNEmbedding, FCEmbedding, EmbeddedNSF
from . import utils, synthetic, dist, retrieval

__all__ = ["NSFFlow", "build_nsf", "posterior_nn", "BoxUniform", "DirectPosterior", "SNPE_C",
           "NSFDensityEstimator", "FCEmbedding", "CNNEmbedding", "EmbeddedNSF", "utils", "synthetic", "dist", "retrieval", "build"]


def build(verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree (sm_100a)."""
    return _lib.build(verbose)
