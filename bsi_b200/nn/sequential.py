"""nn.Sequential that forwards only the keyword arguments each child accepts (reference bsi/nn/sequential.py:6-35)."""

import inspect

from torch import nn


class KwargsSequential(nn.Sequential):
    def __init__(self, *modules):
        super().__init__(*modules)
        self._accepts = []
        for m in modules:
            params = inspect.signature(m.forward).parameters
            takes_all = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())
            self._accepts.append(None if takes_all else frozenset(params))

    def forward(self, input, *args, **kwargs):
        for module, names in zip(self, self._accepts):
            passed = kwargs if names is None else {k: v for k, v in kwargs.items() if k in names}
            input = module(input, *args, **passed)
        return input
