import torch


class ZeroMean(torch.nn.Module):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype)
