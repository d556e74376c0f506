"""``from sbi import utils as utils`` stand-in (modules.py:3): only what the FlopPITy path touches."""
from .inference import BoxUniform
from .flow import posterior_nn

__all__ = ["BoxUniform", "posterior_nn"]
