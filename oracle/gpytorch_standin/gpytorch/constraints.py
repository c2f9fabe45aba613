"""Stand-in for gpytorch.constraints (softplus-transformed bounds)."""
import torch
from torch.nn import functional as F


def inv_softplus(y):
    y = torch.as_tensor(y)
    return y + torch.log(-torch.expm1(-y))


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound)))


class GreaterThan(Interval):
    """transform(raw) = softplus(raw) + lower_bound  (gpytorch.constraints.GreaterThan default transform)."""

    def __init__(self, lower_bound):
        super().__init__(lower_bound, float("inf"))

    def transform(self, raw):
        return F.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return inv_softplus(torch.as_tensor(value) - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
