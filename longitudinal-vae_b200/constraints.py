"""Parameter constraints with GPyTorch's semantics and state_dict keys (softplus transform plus a lower bound).
Used by the drop-in kernel modules and likelihood (LVAE.py:183-184 passes `GreaterThan(1e-8)`)."""
import torch
from torch.nn import functional as F


def inv_softplus(y):
    y = torch.as_tensor(y)
    return y + torch.log(-torch.expm1(-y))


class GreaterThan(torch.nn.Module):
    def __init__(self, lower_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(float("inf")))

    def transform(self, raw):
        return F.softplus(raw) + self.lower_bound

    def lower_float(self):
        """The lower bound as a Python float, read from the buffer once (no device sync on the hot path)."""
        v = self.__dict__.get("_lower_f")
        if v is None:
            v = float(self.lower_bound)
            self.__dict__["_lower_f"] = v
        return v

    def _load_from_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_lower_f", None)        # a checkpoint may carry a different bound: drop the cached float
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_lower_f", None)
        return super()._apply(fn, *args, **kwargs)

    def inverse_transform(self, value):
        return inv_softplus(torch.as_tensor(value) - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
