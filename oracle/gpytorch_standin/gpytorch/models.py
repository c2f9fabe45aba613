import torch


class ExactGP(torch.nn.Module):
    """state_dict container only (GP_def.py:8-21 uses it that way; LVAE.py:195,215,355)."""

    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.likelihood = likelihood
