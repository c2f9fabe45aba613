"""Flattened additive-kernel structure shared by every op of the path.

The reference describes a kernel by six Python lists (parse_model_args.py:74-79) which `kernel_gen.py` turns into a tree
of Scale/Product/leaf modules.  The CUDA ops take the flat form declared in include/lvae_b200.h: one row per additive
component, `outputscale * prod(cat/bin masks) * SE(one column)`, K0 components first, then K1 (the id-covariate ones).
"""
from dataclasses import dataclass

import numpy as np
import torch

from ._lib import BIN, CAT, MAX_COMPONENTS, MAX_MASKS, SPEC_STRIDE


@dataclass
class KernelStructure:
    table: np.ndarray      # int32 [n_comp0+n_comp1, SPEC_STRIDE]
    n_comp0: int
    n_comp1: int
    n_ls: int

    @property
    def n_comp(self):
        return self.n_comp0 + self.n_comp1

    def key(self):
        return (self.table.tobytes(), self.n_comp0, self.n_comp1, self.n_ls)


class Raw:
    """A constrained hyper-parameter that has not been evaluated yet: value = transform(param).
    kind "softplus": softplus(param) + lower (GPyTorch constraints); kind "bounded": exp(lower + softplus(param - lower))
    (GP_model.py:16-18).  build_structure evaluates ALL Raw entries of a kernel with one stack + one transform instead of a
    handful of element-wise launches per parameter (the step at the reference's default batch size is launch-bound)."""
    __slots__ = ("param", "kind", "lower")

    def __init__(self, param, kind, lower):
        self.param, self.kind, self.lower = param, kind, float(lower)

    def numel(self):
        return self.param.numel()

    def value(self):
        p = self.param
        if self.kind == "softplus":
            return torch.nn.functional.softplus(p) + self.lower
        return torch.exp(self.lower + torch.nn.functional.softplus(p - self.lower))


def _val(t):
    return t.value() if isinstance(t, Raw) else t


class FlatComponent:
    """outputscale (tensor [L], Raw, or None == 1) times leaf factors [(kind, dim, lengthscale tensor [L] | Raw | None), ...]."""

    def __init__(self, outputscale, factors):
        self.outputscale = outputscale
        self.factors = factors

    def times(self, other):
        if self.outputscale is None:
            os_ = other.outputscale
        elif other.outputscale is None:
            os_ = self.outputscale
        else:
            os_ = _val(self.outputscale).reshape(-1) * _val(other.outputscale).reshape(-1)
        return FlatComponent(os_, self.factors + other.factors)


def _rows(components):
    """Spec rows + the list of lengthscale tensors, for a list of FlatComponents."""
    rows, ls_list = [], []
    for comp in components:
        row = [-1, 0, 0] + [0] * (SPEC_STRIDE - 3)
        n_mask = 0
        for kind, dim, ls in comp.factors:
            if kind == 'rbf':
                if row[0] >= 0:
                    raise ValueError("lvae_b200: a component may hold at most one squared-exponential factor")
                row[0], row[1] = int(dim), len(ls_list)
                ls_list.append(ls)
            else:
                if n_mask == MAX_MASKS:
                    raise ValueError(f"lvae_b200: more than {MAX_MASKS} categorical/binary factors in one component")
                row[3 + 2 * n_mask] = CAT if kind == 'cat' else BIN
                row[4 + 2 * n_mask] = int(dim)
                n_mask += 1
        row[2] = n_mask
        rows.append(row)
    return rows, ls_list


_LOWER_CACHE = {}


def _pack_raws(entries, L, dtype, device):
    """[n, L] table of the values of `entries` (all Raw, one kind, parameters of 1 or L elements in `dtype` on `device`) with
    one stack and one transform; None if the entries do not qualify."""
    if not entries or not all(isinstance(e, Raw) for e in entries):
        return None
    kind = entries[0].kind
    for e in entries:
        p = e.param
        if e.kind != kind or p.dtype != dtype or p.numel() not in (1, L):
            return None
        if device is not None:
            d = torch.device(device)
            if p.device.type != d.type or (d.index is not None and p.device.index != d.index):
                return None
    rows = [e.param.reshape(-1) if e.param.numel() == L else e.param.reshape(-1).expand(L) for e in entries]
    raw = torch.stack(rows)
    lowers = tuple(e.lower for e in entries)
    if any(lo != 0.0 for lo in lowers) or kind == "bounded":
        key = (lowers, str(raw.device), dtype)
        lo = _LOWER_CACHE.get(key)
        if lo is None:
            lo = torch.tensor(lowers, dtype=dtype, device=raw.device).reshape(-1, 1)
            _LOWER_CACHE[key] = lo
    if kind == "softplus":
        out = torch.nn.functional.softplus(raw)
        return out + lo if any(x != 0.0 for x in lowers) else out
    return torch.exp(lo + torch.nn.functional.softplus(raw - lo))


def build_structure(comps0, comps1, L, dtype=torch.float64, device=None, extra=None):
    """(KernelStructure, lengthscale [n_ls,L], outputscale [n_comp,L]) from two lists of FlatComponents.
    The returned tensors are differentiable functions of the module parameters (softplus etc. stay in PyTorch).
    extra: optional list of further Raw / tensor entries (e.g. the likelihood noise); their [len(extra), L] table is
    returned as a fourth value, evaluated in the same packed transform when possible."""
    comps = list(comps0) + list(comps1)
    if len(comps) > MAX_COMPONENTS:
        raise ValueError(f"lvae_b200: more than {MAX_COMPONENTS} additive components")
    rows, ls_list = _rows(comps)
    table = np.asarray(rows, dtype=np.int32).reshape(len(comps), SPEC_STRIDE)
    if len(ls_list) > MAX_COMPONENTS:
        raise ValueError("lvae_b200: too many lengthscales")

    st = KernelStructure(table=table, n_comp0=len(comps0), n_comp1=len(comps1), n_ls=len(ls_list))
    ex = list(extra) if extra else []
    packed = _pack_raws(ls_list + [c.outputscale for c in comps] + ex, L, dtype, device)
    if packed is not None:                       # one table, three views: [lengthscales | outputscales | extra]
        n_ls, n_c = len(ls_list), len(comps)
        out = (st, packed[:n_ls], packed[n_ls:n_ls + n_c])
        return out + (packed[n_ls + n_c:],) if extra is not None else out

    def as_row(t):
        t = torch.as_tensor(_val(t), dtype=dtype, device=device).reshape(-1)
        return t.expand(L) if t.numel() == 1 else t

    ones = torch.ones(L, dtype=dtype, device=device)
    os_rows = [ones if c.outputscale is None else as_row(c.outputscale) for c in comps]
    ls_rows = [as_row(t) for t in ls_list]
    outputscale = torch.stack(os_rows) if os_rows else torch.zeros(0, L, dtype=dtype, device=device)
    lengthscale = torch.stack(ls_rows) if ls_rows else torch.zeros(0, L, dtype=dtype, device=device)
    out = (st, lengthscale.to(dtype), outputscale.to(dtype))
    if extra is not None:
        ex_rows = [as_row(t) for t in ex]
        out = out + (torch.stack(ex_rows).to(dtype) if ex_rows else torch.zeros(0, L, dtype=dtype, device=device),)
    return out


def flatten(module):
    """FlatComponents of any kernel module of this package (or a duck-typed GPyTorch / GP_model tree)."""
    if hasattr(module, "_flat_components"):
        return module._flat_components()
    # duck typing for foreign trees with the same attribute names (gpytorch.kernels.*)
    name = type(module).__name__
    if name == "AdditiveKernel":
        return [c for k in module.kernels for c in flatten(k)]
    if name == "ProductKernel":
        out = None
        for k in module.kernels:
            cs = flatten(k)
            out = cs if out is None else [a.times(b) for a in out for b in cs]
        return out
    if name == "ScaleKernel":
        return [FlatComponent(module.outputscale.reshape(-1), []).times(c) for c in flatten(module.base_kernel)]
    dim = module.active_dims
    dim = int(dim.reshape(-1)[0]) if torch.is_tensor(dim) else int(dim)
    if name == "RBFKernel":
        return [FlatComponent(None, [('rbf', dim, module.lengthscale.reshape(-1))])]
    if name == "CatKernel":
        return [FlatComponent(None, [('cat', dim, None)])]
    if name == "BinKernel":
        return [FlatComponent(None, [('bin', dim, None)])]
    raise TypeError(f"lvae_b200: cannot flatten kernel module of type {name}")


def latent_count(components, default=1):
    n = default
    for c in components:
        for t in [c.outputscale] + [f[2] for f in c.factors]:
            if t is not None and (torch.is_tensor(t) or isinstance(t, Raw)) and t.numel() > 1:
                n = max(n, t.numel())
    return n
