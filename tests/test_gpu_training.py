"""GPU: the drop-in `hensman_training` loop (training.py:15-237) on a tiny stock-PyTorch VAE and synthetic Health-MNIST-
shaped data, against the same loop written with the oracle's CPU functions — same numpy/torch seeds, same batches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class TinyVAE(torch.nn.Module):
    """Stand-in for ConvVAE (VAE.py:16-162): encode -> (mu, log_var), decode, masked-MSE loss_function."""

    def __init__(self, dim, L):
        super().__init__()
        self.enc = torch.nn.Linear(dim, 2 * L)
        self.dec = torch.nn.Linear(L, dim)
        self.L = L

    def forward(self, x):
        h = self.enc(x.reshape(x.shape[0], -1))
        mu, log_var = h[:, :self.L], -1.0 + 0.1 * h[:, self.L:]
        return self.dec(mu).reshape(x.shape), mu, log_var

    def loss_function(self, recon, x, mask):
        se = ((recon - x) ** 2 * mask).reshape(x.shape[0], -1).sum(1)
        return se, 0.5 * se


class SynthDataset(torch.utils.data.Dataset):
    def __init__(self, x, D, seed):
        g = torch.Generator().manual_seed(seed)
        self.x = x.float()
        self.img = torch.rand(x.shape[0], D, generator=g)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {'digit': self.img[i], 'label': self.x[i], 'idx': i, 'mask': torch.ones_like(self.img[i])}


def _oracle_loop(model, ds, k0, k1, noise, L, m, H, z, P, T, spb, epochs, weight, lr_ng, perms):
    """Same loop on CPU with the oracle; hyper-parameters are optimised through raw = softplus^-1(value), as the modules do."""
    import lvae_oracle as orc
    from lvae_b200.constraints import inv_softplus
    raws = []
    for c in k0 + k1:
        raws.append(inv_softplus(c.outputscale.detach()).clone().requires_grad_(True))
        for k in sorted(c.lengthscales):
            raws.append(inv_softplus(c.lengthscales[k].detach()).clone().requires_grad_(True))

    def refresh():
        it = iter(raws)
        for c in k0 + k1:
            c.outputscale = torch.nn.functional.softplus(next(it))
            for k in sorted(c.lengthscales):
                c.lengthscales[k] = torch.nn.functional.softplus(next(it))
    opt = torch.optim.Adam(list(model.parameters()) + raws, lr=1e-3)
    kld_curve = []
    for ep in range(epochs):
        batches = orc.fixed_T_batches(perms[ep], T, spb)
        tot = 0.0
        for rows in batches:
            opt.zero_grad()
            refresh()
            data = torch.stack([ds[i]['digit'] for i in rows]).double()
            x = torch.stack([ds[i]['label'] for i in rows]).double()
            recon, mu, lv = model(data)
            rl, _ = model.loss_function(recon, data, torch.ones_like(data))
            Pb = len(rows) // T
            kld, gm, gH = orc.kld_fixed_T(k0, k1, noise, L, m, H, x, mu, lv, z, P, Pb, T, True, 1e-6)
            loss = rl.sum() * P / Pb + weight * kld / L
            loss.backward()
            opt.step()
            m, H = orc.ng_step(m, H, gm.detach(), gH.detach(), lr_ng)
            tot += (kld / L).item() / len(batches)
        kld_curve.append(tot)
    return np.array(kld_curve), m, H


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_hensman_training_matches_oracle_loop(cuda_graph):
    import copy
    import lvae_oracle as orc
    from helpers import build_modules, rel
    from lvae_b200 import synth
    from lvae_b200.training import hensman_training
    P, T, L, M, D, spb, epochs = 8, 20, 3, 12, 16, 3, 2
    b = synth.make_batch("cfg2", P=P, L=L, M=M)
    ds = SynthDataset(b.x, D, seed=5)
    torch.manual_seed(0)
    model_cpu = TinyVAE(D, L).double()
    model_gpu = copy.deepcopy(model_cpu).cuda()
    k0, k1 = orc.parse_kernel_lists(L, **b.lists, id_covariate=2)
    noise = torch.ones(L, dtype=torch.float64)
    ls = torch.stack([c.lengthscales[k].detach() for c in k0 + k1 for k in sorted(c.lengthscales)])
    os_ = torch.stack([c.outputscale.detach() for c in k0 + k1])
    # permutations the SubjectSampler will draw
    np.random.seed(42)
    perms = []
    for _ in range(epochs):
        r = np.arange(P)
        np.random.shuffle(r)
        perms.append(r.copy())
    ref_curve, m_ref, H_ref = _oracle_loop(model_cpu, ds, k0, k1, noise, L, b.m.clone(), b.H.clone(), b.z, P, T, spb, epochs,
                                           0.15, 0.01, perms)
    cm0, cm1, lik = build_modules(b.lists, L, ls.numpy(), os_.numpy(), noise.numpy())
    lik.noise_covar.raw_noise.requires_grad_(False)
    opt = torch.optim.Adam(list(model_gpu.parameters()) + list(cm0.parameters()) + list(cm1.parameters()), lr=1e-3)
    np.random.seed(42)
    out = hensman_training(model_gpu, 'conv', epochs, ds, opt, 'GPapprox_closed', 1, L, cm0, cm1, lik, b.m.cuda(), b.H.cuda(),
                           b.z.cuda(), P, T, False, 6, 0.15, 2, 'mse', natural_gradient=True, natural_gradient_lr=0.01,
                           subjects_per_batch=spb, num_workers=0, verbose=False, cuda_graph=cuda_graph)
    # identical batches, initial weights and optimiser state on both sides: the per-epoch GP loss and the final (m, H) agree
    assert out[4].shape == (epochs,)
    assert np.abs(out[4] - ref_curve).max() <= 1e-6 * np.abs(ref_curve).max()
    assert rel(out[5].cpu(), m_ref) < 1e-6 and rel(out[6].cpu(), H_ref) < 1e-6
