"""Parameter container with the reference's MLP key layout (reference bsi/nn/mlp.py:6-40):
`0`, `2`, ... are the Linear layers, odd indices the activations."""

from typing import Callable

from torch import nn


class MLP(nn.Sequential):
    def __init__(self, in_features: int, out_features: int, *, hidden_features, hidden_layers: int | None = None,
                 actfn: Callable[[], nn.Module] = nn.Identity):
        if isinstance(hidden_features, int):
            assert hidden_layers is not None and hidden_layers >= 0
            hidden_features = [hidden_features] * hidden_layers
        elif hidden_layers is not None:
            assert len(hidden_features) == hidden_layers
        widths = [in_features, *hidden_features, out_features]
        layers: list[nn.Module] = []
        for i, (a, b) in enumerate(zip(widths[:-1], widths[1:])):
            layers.append(nn.Linear(a, b))
            if i < len(widths) - 2:
                layers.append(actfn())
        self.in_features, self.out_features = in_features, out_features
        self.hidden_features, self.hidden_layers, self.actfn = list(hidden_features), len(hidden_features), actfn
        super().__init__(*layers)
