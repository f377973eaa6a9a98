from .cdae import CDAE
from .mf import MatrixFactorization
from .ngcf import NGCF

__all__ = ["MatrixFactorization", "NGCF", "CDAE"]
