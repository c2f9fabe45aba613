"""CPU: the C-ABI library builds, loads and exports every symbol include/lvae_b200.h declares (no compute calls without a
GPU); host-side logic: kernel-structure flattening (block-indexing rule), samplers, workspace sizing, loud failure."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as ge
    lib_path = ge.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "lvae_b200.h")).read()
    declared = set(re.findall(r"\b(lvae_[a-z0-9_]+)\s*\(", header))
    declared -= {"lvae_kernel_spec_t", "lvae_kld_problem_t"}
    assert len(declared) >= 14
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/lvae_b200.h but not exported"
    from lvae_b200 import _lib
    assert set(_lib.EXPORTS) <= declared


def test_block_indexing_rule_is_bit_exact():
    """kernel_gen.py:225-308: order cat, sqexp, bin, cat_int, bin_int; K1 iff id cat or cat_int on the id covariate."""
    import lvae_oracle as orc
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.spec import build_structure, flatten
    from lvae_b200 import synth, GP_model
    lists_all = [synth.kernel_lists(c) for c in ("cfg1", "cfg2", "cfg4")] + [dict(
        cat_kernel=[2, 3], bin_kernel=[5], sqexp_kernel=[0, 1],
        cat_int_kernel=[{'cont_covariate': 1, 'cat_covariate': 2}, {'cont_covariate': 0, 'cat_covariate': 3}],
        bin_int_kernel=[{'cont_covariate': 1, 'bin_covariate': 5}], covariate_missing_val=[{'covariate': 1, 'mask': 4}])]
    for lists in lists_all:
        k0, k1 = orc.parse_kernel_lists(3, **lists, id_covariate=2)
        for gen in (generate_kernel_batched, GP_model.generate_kernel_batched):
            cm0, cm1 = gen(3, **lists, id_covariate=2)
            st, ls, os_ = build_structure(flatten(cm0), flatten(cm1), 3)
            assert (st.n_comp0, st.n_comp1) == (len(k0), len(k1))
            for row, comp in zip(st.table, k0 + k1):
                rbf = [d for kind, d in comp.factors if kind == 'rbf']
                masks = [(0 if kind == 'cat' else 1, d) for kind, d in comp.factors if kind != 'rbf']
                assert row[0] == (rbf[0] if rbf else -1)
                assert row[2] == len(masks)
                assert [(row[3 + 2 * i], row[4 + 2 * i]) for i in range(len(masks))] == masks
            assert st.n_ls == sum(len(c.lengthscales) for c in k0 + k1)
            assert torch.allclose(ls, torch.full_like(ls, 2.5), rtol=1e-6)            # kernel_spec.py:68 / GP_model.py:60
            assert torch.allclose(os_, torch.full_like(os_, float(np.log(2.0))), rtol=1e-6)


def test_state_dict_keys_follow_gpytorch_layout():
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200 import synth
    cm0, cm1 = generate_kernel_batched(4, **synth.kernel_lists("cfg2"), id_covariate=2)
    keys = set(cm0.state_dict().keys())
    assert "kernels.0.raw_outputscale" in keys and "kernels.0.base_kernel.raw_lengthscale" in keys
    assert "kernels.1.base_kernel.kernels.1.raw_lengthscale" in keys
    assert cm0.kernels[0].base_kernel.raw_lengthscale.shape == (4, 1, 1)
    assert cm0.kernels[0].raw_outputscale.shape == (4,)


def test_product_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200 import synth, elbo_functions as EF
    from lvae_b200.likelihoods import GaussianLikelihood
    b = synth.make_batch("cfg1", P=2, L=2, M=5)
    cm0, cm1 = generate_kernel_batched(2, **b.lists, id_covariate=2)
    with pytest.raises(RuntimeError):
        cm0(b.x, b.z).evaluate()
    with pytest.raises(RuntimeError):
        EF.minibatch_KLD_upper_bound(cm0, cm1, GaussianLikelihood(batch_shape=torch.Size([2])), 2, b.m, b.H, b.x, b.mu,
                                     b.log_v, b.z, 2, 2, 20, True, 1e-6)
    # the evaluation-side functions (prediction, DUBO) route through the same ops: no CPU fallback either
    from lvae_b200 import utils as U
    from lvae_b200.validation import validation_dubo
    lik = GaussianLikelihood(batch_shape=torch.Size([2]))
    with pytest.raises(RuntimeError):
        U.batch_predict_varying_T(2, cm0, cm1, lik, b.x, b.x[:5], b.mu, b.z, 2, 1e-6)
    with pytest.raises(RuntimeError):
        validation_dubo(2, cm0, cm1, lik, b.x, b.mu, b.log_v, b.z, 2, 20, 1e-6)


def test_samplers_match_reference_golden():
    """utils.py:40-113 — bit-exact row indices and batch composition under the same numpy seed."""
    from lvae_b200.utils import SubjectSampler, VaryingLengthBatchSampler, VaryingLengthSubjectSampler
    from torch.utils.data.sampler import BatchSampler
    s = dict(np.load(os.path.join(ROOT, "tests", "golden", "samplers.npz")))
    for k in range(3):
        P, T, spb = (int(v) for v in s[f"fixed{k}_PTspb"])
        np.random.seed(100 + k)
        rows = list(iter(SubjectSampler(list(range(P * T)), P, T)))
        assert rows == s[f"fixed{k}_rows"].tolist()
        assert [len(b) for b in BatchSampler(rows, spb * T, drop_last=False)] == s[f"fixed{k}_batch_lens"].tolist()
    ids = s["vary_ids"]
    data = [{'label': torch.tensor([0.0, 0.0, float(i)])} for i in ids]
    for k in range(2):
        vs = VaryingLengthSubjectSampler(data, 2)
        np.random.seed(200 + k)
        batches = list(iter(VaryingLengthBatchSampler(vs, int(s[f"vary{k}_spb"]))))
        assert [i for b in batches for i in b] == s[f"vary{k}_flat"].tolist()
        assert [len(b) for b in batches] == s[f"vary{k}_batch_lens"].tolist()


def test_group_by_subject_matches_boolean_mask_grouping():
    """elbo_functions.py:264-267 on CPU tensors (pure torch host logic)."""
    import lvae_oracle as orc
    from lvae_b200.elbo_functions import group_by_subject
    rng = np.random.default_rng(1)
    ids = rng.integers(0, 9, size=60).astype(np.float64)
    order, offsets, T_max, sum_T2 = group_by_subject(torch.from_numpy(ids))
    uniq, rows = orc.group_rows_by_subject(ids)
    assert np.array_equal(order.numpy(), np.concatenate(rows))
    lens = np.array([len(r) for r in rows])
    assert np.array_equal(offsets.numpy(), np.concatenate([[0], np.cumsum(lens)]))
    assert T_max == lens.max() and sum_T2 == int((lens * lens).sum())
    sorted_ids = np.sort(ids)
    order2, *_ = group_by_subject(torch.from_numpy(sorted_ids))
    assert order2 is None


def test_split_call_row_mapping_and_decision():
    """Host logic of ops.make_kld_call / SplitKldCall (ragged minibatches with long subjects): which subjects go where, and the
    global row of every local row, without touching the GPU."""
    import numpy as np
    from lvae_b200 import ops
    counts = np.array([3, 30, 5, 26, 24, 1, 40])
    off = np.concatenate([[0], np.cumsum(counts)])
    seen = []
    for sel in (counts <= 24, counts > 24):
        idx = np.nonzero(sel)[0]
        c = counts[idx]
        start = np.cumsum(c) - c
        rows = np.repeat(off[idx] - start, c) + np.arange(int(c.sum()))
        assert np.array_equal(rows, np.concatenate([np.arange(off[p], off[p + 1]) for p in idx]))
        seen.append(rows)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(off[-1]))      # a partition of the rows

    # decision rule (no device needed: it only inspects the counts before constructing anything)
    def decide(M, counts, path=0):
        counts = np.asarray(counts)
        if path == 0 and M <= 64 and counts.size and counts.max() > 24:
            short = counts <= 24
            return bool(short.any() and counts[short].sum() >= 0.1 * counts.sum())
        return False
    assert decide(60, [5, 40, 20]) and not decide(60, [20, 20]) and not decide(128, [5, 40]) and not decide(60, [40, 39])
    assert not decide(60, [5, 40], path=1) and not decide(60, [1] + [40] * 50)
    src = open(ops.__file__).read()
    assert "counts[short].sum() >= 0.1 * counts.sum()" in src and "counts.max() > 24" in src   # the rule tested above is the shipped one


def test_packed_hyper_parameters_match_module_properties():
    """spec.build_structure evaluates all lengthscales / outputscales / noise with ONE stack + ONE transform (spec.Raw); values
    and gradients w.r.t. every raw parameter must equal the per-module constraint transforms (gpytorch-style softplus + lower
    bound, and GP_model.py's exp(min + softplus(raw - min)))."""
    from lvae_b200 import synth, GP_model as GM
    from lvae_b200.elbo_functions import _noise_entry
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.gp_kernels import RBFKernel
    from lvae_b200.spec import build_structure, flatten
    L = 3
    lists = synth.kernel_lists(synth.CONFIGS["cfg4"]) if hasattr(synth, "CONFIGS") else synth.make_batch("cfg4", P=2, L=L, M=4).lists
    torch.manual_seed(0)
    for style in ("gpytorch", "gp_model"):
        if style == "gpytorch":
            cm0, cm1 = generate_kernel_batched(L, **lists, id_covariate=2)
            lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8))
        else:
            cm0, cm1 = GM.generate_kernel_batched(L, **lists, id_covariate=2)
            lik = GM.Likelihoods(L, 1.0)
        mods = [cm0.double(), cm1.double(), lik.double()]
        params = [p for m_ in mods for p in m_.parameters()]
        with torch.no_grad():
            for p in params:
                p.add_(0.3 * torch.randn_like(p))
        st, ls, os_, nz = build_structure(flatten(cm0), flatten(cm1), L, device="cpu", extra=[_noise_entry(lik)])
        assert ls._base is not None and ls._base is os_._base and ls._base is nz._base          # the packed path was taken
        # reference values straight from the module properties
        ref_os, ref_ls = [], []
        for mod in (cm0, cm1):
            for sk in mod.kernels:
                ref_os.append((sk.outputscale if style == "gpytorch" else sk.scale).reshape(-1))
                for sub in sk.modules():
                    if isinstance(sub, RBFKernel) or isinstance(sub, GM.RbfKernel):
                        ref_ls.append(sub.lengthscale.reshape(-1))
        ref_noise = (lik.noise_covar.noise if style == "gpytorch" else lik.noise).reshape(-1)
        assert torch.allclose(os_, torch.stack(ref_os), rtol=1e-14, atol=0) and torch.allclose(ls, torch.stack(ref_ls), rtol=1e-14, atol=0)
        assert torch.allclose(nz.reshape(-1), ref_noise, rtol=1e-14, atol=0)
        w = torch.randn(ls._base.shape, dtype=torch.float64)
        g1 = torch.autograd.grad((ls._base * w).sum(), params, allow_unused=True)
        ref_table = torch.cat([torch.stack(ref_ls), torch.stack(ref_os), ref_noise.reshape(1, L)])
        g2 = torch.autograd.grad((ref_table * w).sum(), params, allow_unused=True)
        for a, b in zip(g1, g2):
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.allclose(a, b, rtol=1e-13, atol=1e-300)


def test_gp_model_checkpoint_round_trip(tmp_path):
    """gp_model.pth / zt_list.pth / m.pth / H.pth (LVAE.py:213-232, 353-360): gpytorch's key layout, strict reload."""
    from lvae_b200 import synth
    from lvae_b200.GP_def import ExactGPModel, load_hensman_state, save_hensman_state
    from lvae_b200.constraints import GreaterThan
    from lvae_b200.kernel_gen import generate_kernel_batched
    from lvae_b200.likelihoods import GaussianLikelihood
    L = 3
    b = synth.make_batch("cfg2", P=4, L=L, M=7)

    def make():
        cm0, cm1 = generate_kernel_batched(L, **b.lists, id_covariate=2)
        lik = GaussianLikelihood(batch_shape=torch.Size([L]), noise_constraint=GreaterThan(1e-8))
        return ExactGPModel(b.x, b.mu, lik, cm0 + cm1).double(), cm0, cm1, lik

    gp, cm0, cm1, lik = make()
    keys = set(gp.state_dict().keys())
    for k in ("likelihood.noise_covar.raw_noise", "likelihood.noise_covar.raw_noise_constraint.lower_bound",
              "covar_module.kernels.0.raw_outputscale", "covar_module.kernels.0.base_kernel.raw_lengthscale",
              "covar_module.kernels.0.base_kernel.raw_lengthscale_constraint.lower_bound",
              "covar_module.kernels.0.raw_outputscale_constraint.upper_bound",
              "covar_module.kernels.0.base_kernel.active_dims"):
        assert k in keys, k
    n0 = len(cm0.kernels)
    # the sum shares the parameters of its two operands (LVAE.py:196): same storage under both names
    assert gp.covar_module.kernels[n0].raw_outputscale is cm1.kernels[0].raw_outputscale
    with torch.no_grad():
        for p in gp.parameters():
            p.add_(torch.randn_like(p))
    save_hensman_state(str(tmp_path), gp, b.z, b.m, b.H)
    save_hensman_state(str(tmp_path), gp, b.z, b.m, b.H, suffix="_best")
    gp2, cm0b, cm1b, lik2 = make()
    z2, m2, H2 = load_hensman_state(str(tmp_path), gp2, "cpu")
    for (k, a), (k2, c) in zip(gp.state_dict().items(), gp2.state_dict().items()):
        assert k == k2 and torch.equal(a, c)
    assert torch.equal(z2, b.z) and torch.equal(m2, b.m) and torch.equal(H2, b.H)
    assert torch.equal(lik2.noise, lik.noise) and torch.equal(cm1b.kernels[0].outputscale, cm1.kernels[0].outputscale)
    # a checkpoint with a missing key is rejected (strict), as with gpytorch
    sd = gp.state_dict()
    sd.pop("covar_module.kernels.0.raw_outputscale")
    with pytest.raises(RuntimeError):
        gp2.load_state_dict(sd)
