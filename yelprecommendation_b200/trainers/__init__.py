from .cdae_trainer import CDAETrainer
from .chunked_evaluator import ChunkedTopKEvaluator
from .mf_trainer import MFTrainer
from .ngcf_trainer import NGCFTrainer

__all__ = ["MFTrainer", "NGCFTrainer", "CDAETrainer", "ChunkedTopKEvaluator"]
