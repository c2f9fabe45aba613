"""Drop-in for the sampler half of the reference's utils.py (utils.py:9-113): subject-wise batching for the Hensman loop.

Semantics kept bit-exact under the same numpy RNG state (tests/golden/samplers.npz):
  SubjectSampler                 np.random.shuffle(arange(P)), then rows T*s .. T*s+T-1 of each subject in that order
  VaryingLengthSubjectSampler    ids = int(label[id_covariate]); a subject's rows run from the first occurrence of its id to
                                 the first occurrence of the next new id (rows assumed contiguous); shuffled subjects
  VaryingLengthBatchSampler      a batch closes when `batch_size` distinct subjects have been collected
  HensmanDataLoader              one persistent iterator over an endlessly repeating batch sampler; len = batches/epoch
"""
import numpy as np
import torch
from torch.utils.data.sampler import BatchSampler, Sampler


class _RepeatSampler:
    """Wraps a (batch) sampler so that iterating never ends."""

    def __init__(self, sampler):
        self.sampler = sampler

    def __iter__(self):
        while True:
            for item in self.sampler:
                yield item


class HensmanDataLoader(torch.utils.data.dataloader.DataLoader):
    def __init__(self, dataset, batch_sampler, num_workers):
        super().__init__(dataset, batch_sampler=_RepeatSampler(batch_sampler), num_workers=num_workers)
        self.iterator = super().__iter__()

    def __len__(self):
        return len(self.batch_sampler.sampler)

    def __iter__(self):
        for _ in range(len(self)):
            yield next(self.iterator)


class SubjectSampler(Sampler):
    def __init__(self, data_source, P, T):
        self.data_source, self.P, self.T = data_source, P, T

    def __iter__(self):
        order = np.arange(self.P)
        np.random.shuffle(order)
        rows = (order[:, None] * self.T + np.arange(self.T)[None, :]).reshape(-1)
        return iter(rows.tolist())

    def __len__(self):
        return len(self.data_source)


class VaryingLengthSubjectSampler(Sampler):
    def __init__(self, data_source, id_covariate):
        self.data_source, self.id_covariate = data_source, id_covariate
        ids = [int(sample['label'][id_covariate].item()) for sample in data_source]
        first = {}
        for row, v in enumerate(ids):
            first.setdefault(v, row)
        self.P = len(first)
        self.start_indices = list(first.values())                       # order of first appearance
        self.end_indices = self.start_indices[1:] + [len(data_source)]

    def __iter__(self):
        order = np.arange(self.P)
        np.random.shuffle(order)
        return iter([(row, int(s)) for s in order for row in range(self.start_indices[s], self.end_indices[s])])

    def __len__(self):
        return self.P


class VaryingLengthBatchSampler(BatchSampler):
    def __init__(self, sampler, batch_size):
        super().__init__(sampler, batch_size, False)
        assert isinstance(sampler, VaryingLengthSubjectSampler)
        self.sampler, self.batch_size = sampler, batch_size

    def __iter__(self):
        rows, members = [], set()
        for row, subject in self.sampler:
            if subject not in members:
                if len(members) == self.batch_size:
                    yield rows
                    rows, members = [], set()
                members.add(subject)
            rows.append(row)
        yield rows
