"""Mirror of the reference's models/base_model.py:6-18 (abstract nn.Module with an activation factory)."""
from abc import ABC, abstractmethod

import torch.nn as nn


class BaseModel(nn.Module, ABC):
    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)

    def _activation_module(self, function_name: str) -> nn.Module:
        return {"sigmoid": nn.Sigmoid, "identity": nn.Identity}[function_name]() \
            if function_name in ("sigmoid", "identity") else None

    @abstractmethod
    def _init_weights(self):
        ...
