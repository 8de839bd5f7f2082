"""`MultipleOutputLoss2`: deep-supervision wrapper with the interface of the reference class
(nnunet/training/loss_functions/deep_supervision.py:30-54).  Semantics kept: the highest-resolution scale is always
evaluated (even with weight 0), any further scale only when its weight is non-zero, result = weighted sum."""
from torch import nn


class MultipleOutputLoss2(nn.Module):
    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.loss = loss
        self.weight_factors = weight_factors

    def _scale_weights(self, n_scales):
        if self.weight_factors is None:
            return [1] * n_scales
        return list(self.weight_factors)

    def forward(self, x, y):
        for name, seq in (("x", x), ("y", y)):
            assert isinstance(seq, (tuple, list)), "%s must be either tuple or list" % name
        w = self._scale_weights(len(x))
        total = w[0] * self.loss(x[0], y[0])
        for wi, xi, yi in zip(w[1:len(x)], x[1:], y[1:]):
            if wi == 0:
                continue
            total = total + wi * self.loss(xi, yi)
        return total
