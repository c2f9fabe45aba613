"""CPU: property tests (hypothesis) of the integer / index work on the path — bit-exact against the oracle's restatement of
utils.py:40-113 and elbo_functions.py:264-267 on random shapes, not just the committed golden cases."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st
from torch.utils.data.sampler import BatchSampler

import lvae_oracle as orc
from lvae_b200.utils import HensmanDataLoader, SubjectSampler, VaryingLengthBatchSampler, VaryingLengthSubjectSampler

CFG = dict(max_examples=60, deadline=None)


def _perm(seed, P):
    np.random.seed(seed)
    r = np.arange(P)
    np.random.shuffle(r)
    return r


@settings(**CFG)
@given(P=st.integers(1, 40), T=st.integers(1, 12), spb=st.integers(1, 45), seed=st.integers(0, 2**31 - 1))
def test_fixed_T_sampler_and_batches(P, T, spb, seed):
    perm = _perm(seed, P)
    np.random.seed(seed)
    rows = list(iter(SubjectSampler(list(range(P * T)), P, T)))
    assert rows == orc.subject_sampler_rows(perm, T)
    np.random.seed(seed)
    batches = list(BatchSampler(SubjectSampler(list(range(P * T)), P, T), spb * T, drop_last=False))
    assert batches == orc.fixed_T_batches(perm, T, spb)
    # every batch holds whole subjects; only the last one may be short (training.py:114: P_in_current_batch = N_batch // T)
    assert all(len(b) % T == 0 for b in batches) and all(len(b) == spb * T for b in batches[:-1])
    assert sorted(r for b in batches for r in b) == list(range(P * T))


@settings(**CFG)
@given(lens=st.lists(st.integers(1, 9), min_size=1, max_size=25), spb=st.integers(1, 30), seed=st.integers(0, 2**31 - 1),
       relabel=st.booleans())
def test_varying_T_sampler_and_batches(lens, spb, seed, relabel):
    P = len(lens)
    labels = np.arange(P)
    if relabel:                                   # ids need not be 0..P-1 nor sorted: subjects are numbered by first appearance
        labels = np.random.default_rng(seed).permutation(P) * 3 + 5
    ids = np.repeat(labels, lens)
    data = [{'label': torch.tensor([0.0, float(i), 1.0])} for i in ids]
    vs = VaryingLengthSubjectSampler(data, 1)
    starts, ends = orc.varying_T_index(ids)
    assert vs.start_indices == starts and vs.end_indices == ends and len(vs) == P
    perm = _perm(seed, P)
    np.random.seed(seed)
    batches = list(iter(VaryingLengthBatchSampler(vs, spb)))
    assert batches == orc.varying_T_batches(ids, perm, spb)
    assert sorted(r for b in batches for r in b) == list(range(len(ids)))
    subj_of_row = np.repeat(np.arange(P), lens)
    assert all(len(set(subj_of_row[b].tolist())) == spb for b in batches[:-1])


@settings(**CFG)
@given(ids=st.lists(st.integers(0, 12), min_size=1, max_size=80))
def test_group_by_subject_any_row_order(ids):
    from lvae_b200.elbo_functions import group_by_subject
    ids = np.asarray(ids, dtype=np.float64)
    order, offsets, T_max, sum_T2 = group_by_subject(torch.from_numpy(ids))
    uniq, rows = orc.group_rows_by_subject(ids)
    lens = np.array([len(r) for r in rows])
    want = np.concatenate(rows)
    got = np.arange(len(ids)) if order is None else order.numpy()
    assert np.array_equal(got, want)                      # subjects ascending, rows of a subject in their original order
    assert np.array_equal(offsets.numpy(), np.concatenate([[0], np.cumsum(lens)]))
    assert T_max == lens.max() and sum_T2 == int((lens * lens).sum())
    assert np.array_equal(group_by_subject.last_counts, lens)


def test_hensman_loader_repeats_epochs_without_restarting():
    """utils.py:9-38: len = batches per epoch; one persistent iterator, so a new `for` continues with a new permutation."""
    P, T, spb = 7, 3, 3

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return P * T

        def __getitem__(self, i):
            return {'idx': torch.tensor(i)}

    np.random.seed(3)
    dl = HensmanDataLoader(DS(), batch_sampler=BatchSampler(SubjectSampler(DS(), P, T), spb * T, drop_last=False),
                           num_workers=0)
    assert len(dl) == 3
    epochs = [[b['idx'].tolist() for b in dl] for _ in range(3)]
    np.random.seed(3)
    for ep in epochs:
        r = np.arange(P)
        np.random.shuffle(r)
        assert ep == orc.fixed_T_batches(r, T, spb)


_COV = st.integers(0, 5)


@settings(max_examples=40, deadline=None)
@given(cat=st.lists(_COV, max_size=3, unique=True), bin_=st.lists(_COV, max_size=2, unique=True),
       sq=st.lists(_COV, max_size=3, unique=True),
       cat_int=st.lists(st.tuples(_COV, _COV), max_size=3), bin_int=st.lists(st.tuples(_COV, _COV), max_size=2),
       missing=st.lists(st.tuples(_COV, _COV), max_size=2, unique_by=lambda t: t[0]), id_cov=_COV)
def test_block_indexing_rule_on_random_kernel_lists(cat, bin_, sq, cat_int, bin_int, missing, id_cov):
    """kernel_gen.py:225-308 / GP_model.py:146-236 on random structure lists: K0/K1 assignment, component order, factor order
    and missing-value masks of both generators equal the oracle's restatement, bit for bit."""
    from lvae_b200 import GP_model
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.spec import build_structure, flatten
    lists = dict(cat_kernel=cat, bin_kernel=bin_, sqexp_kernel=sq,
                 cat_int_kernel=[{'cont_covariate': c, 'cat_covariate': k} for c, k in cat_int],
                 bin_int_kernel=[{'cont_covariate': c, 'bin_covariate': k} for c, k in bin_int],
                 covariate_missing_val=[{'covariate': c, 'mask': m} for c, m in missing])
    if id_cov not in cat and not any(k == id_cov for _, k in cat_int):
        lists['cat_kernel'] = cat + [id_cov]     # the reference needs at least one component on either side
    if not (lists['sqexp_kernel'] or bin_ or bin_int or any(d != id_cov for d in lists['cat_kernel'])
            or any(k != id_cov for _, k in cat_int)):
        lists['sqexp_kernel'] = [0]
    k0, k1 = orc.parse_kernel_lists(2, **lists, id_covariate=id_cov)
    assert k0 and k1
    for gen in (generate_kernel_batched, GP_model.generate_kernel_batched):
        cm0, cm1 = gen(2, **lists, id_covariate=id_cov)
        stc, ls, os_ = build_structure(flatten(cm0), flatten(cm1), 2)
        assert (stc.n_comp0, stc.n_comp1) == (len(k0), len(k1))
        assert stc.n_ls == sum(len(c.lengthscales) for c in k0 + k1)
        i_ls = 0
        for row, comp in zip(stc.table, k0 + k1):
            rbf = [d for kind, d in comp.factors if kind == 'rbf']
            masks = [(0 if kind == 'cat' else 1, d) for kind, d in comp.factors if kind != 'rbf']
            assert row[0] == (rbf[0] if rbf else -1) and row[2] == len(masks)
            assert [(row[3 + 2 * i], row[4 + 2 * i]) for i in range(len(masks))] == masks
            if rbf:
                assert row[1] == i_ls
                i_ls += 1


def test_call_pool_hands_out_only_unreferenced_calls():
    """elbo_functions._pooled_call: a prepared call is reused only when nothing refers to it any more; live references
    (an autograd node, the deferred-check queue, a tag on grad_H) get the next caller a different object; at most
    _POOL_SIGNATURES problem shapes are remembered."""
    import lvae_b200.elbo_functions as EF

    class Call:
        pass
    EF.clear_call_pool()
    built = []

    def make():
        built.append(1)                            # count only: a reference kept here would pin the call
        return Call()
    a = EF._pooled_call("k", make)
    b = EF._pooled_call("k", make)                 # `a` still referenced here
    assert a is not b and len(built) == 2
    ida = id(a)
    del a
    c = EF._pooled_call("k", make)                 # the first one is free again
    assert id(c) == ida and len(built) == 2
    holder = [c]                                   # e.g. the deferred-check queue
    del c
    d = EF._pooled_call("k", make)
    assert d is not holder[0] and d is not b and len(built) == 3
    holder.clear()
    for i in range(EF._POOL_SIGNATURES + 3):       # ragged minibatches: new shapes push old ones out
        EF._pooled_call(("shape", i), make)
    assert len(EF._POOL) == EF._POOL_SIGNATURES and "k" not in EF._POOL
    EF.clear_call_pool()
    assert not EF._POOL
