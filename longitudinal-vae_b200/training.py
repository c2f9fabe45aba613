"""Host side of the Hensman minibatch loop (training.py:15-237) over the CUDA GP-prior ops.

`hensman_training` keeps the reference's positional signature and return tuple.  The step body follows
training.py:90-140 line by line; the three library calls on the hot path are `minibatch_KLD_upper_bound[_iter]`,
autograd through it, and `natural_gradient_step` (training.py:129-135 as one launch reusing the H^-1 the bound already
computed).  The periodic validation / test-set MSE / checkpoint block (training.py:150-235) belongs to the reference's
evaluation code (out of scope): it is reduced to an optional `on_validation(epoch, state)` callback.
"""
import numpy as np
import torch
from torch.utils.data.sampler import BatchSampler

from . import ops
from .elbo_functions import minibatch_KLD_upper_bound, minibatch_KLD_upper_bound_iter, subject_counts_of
from .utils import HensmanDataLoader, SubjectSampler, VaryingLengthBatchSampler, VaryingLengthSubjectSampler


_NG_PENDING = []      # (pinned flags, event) of natural-gradient updates whose failure flag has not been read yet


def check_natural_gradient_errors(wait=True):
    """Raise if a natural-gradient update hit a non-positive-definite H^-1 + lr (gH + gH^T) (torch.cholesky raises at that
    point in the reference, training.py:133).  wait=False only looks at updates whose flags have already arrived."""
    while _NG_PENDING:
        flags, ev = _NG_PENDING[0]
        if not wait and not ev.query():
            return
        ev.synchronize()
        _NG_PENDING.pop(0)
        if int(flags[3]) != 0:
            _NG_PENDING.clear()
            raise RuntimeError(f"cholesky: the natural-gradient update of latent {int(flags[3]) - 1} is not positive-definite "
                               "(H^-1 + lr (grad_H + grad_H^T), training.py:131-133)")


def natural_gradient_step(m, H, grad_m, grad_H, natural_gradient_lr, check=False):
    """training.py:129-135 in one launch: iH = H^-1; iH' = iH + lr (gH + gH^T); H <- iH'^-1;
    m <- H (iH m - lr (g_m - 2 gH m)).  Returns detached (m, H).
    A non-positive-definite iH' raises like the reference's torch.cholesky: at once with check=True (one device sync),
    otherwise deferred — the flag travels to pinned host memory without blocking and is looked at by the next call(s) and by
    check_natural_gradient_errors() (hensman_training calls it after the last step)."""
    check_natural_gradient_errors(wait=False)
    lt = getattr(grad_H, "_lvae_latent_tail", None)
    if lt is not None:          # latent-sharded tail: this rank updates its latents, the new (m, H) are all-gathered
        m2, H2 = m.detach().to(torch.float64).contiguous().clone(), H.detach().to(torch.float64).contiguous().clone()
        lt.ng_step(m2, H2, natural_gradient_lr)
        if check and int(lt.ng_info[3].item()) != 0:
            raise RuntimeError("cholesky: the natural-gradient update is not positive-definite (training.py:131-133)")
        return m2.view_as(m), H2
    hinv = None
    tag = getattr(grad_H, "_lvae_hinv", None)          # H^-1 computed by the bound for this very H (same storage, unmodified)
    if tag is not None and tag[1] == H.data_ptr() and tag[2] == H._version and H.dtype == tag[0].dtype:
        hinv = tag[0]
    m2, H2, info = ops.ng_step(m, H, grad_m, grad_H, natural_gradient_lr, Hinv=hinv)
    if check:
        if int(info[3].item()) != 0:
            raise RuntimeError(f"cholesky: the natural-gradient update of latent {int(info[3].item()) - 1} is not "
                               "positive-definite (H^-1 + lr (grad_H + grad_H^T), training.py:131-133)")
    else:
        flags = torch.empty(4, dtype=torch.int32).pin_memory()
        flags.copy_(info, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(H.device))
        _NG_PENDING.append((flags, ev))
        if len(_NG_PENDING) > 4:
            check_natural_gradient_errors(wait=False)
    return m2.detach(), H2.detach()


def hensman_training(nnet_model, type_nnet, epochs, dataset, optimiser, type_KL, num_samples, latent_dim, covar_module0,
                     covar_module1, likelihoods, m, H, zt_list, P, T, varying_T, Q, weight, id_covariate, loss_function,
                     natural_gradient=False, natural_gradient_lr=0.01, subjects_per_batch=20, memory_dbg=False,
                     eps=1e-6, results_path=None, validation_dataset=None, generation_dataset=None,
                     prediction_dataset=None, gp_model=None, csv_file_test_data=None, csv_file_test_label=None,
                     test_mask_file=None, data_source_path=None, num_workers=4, on_validation=None, verbose=True,
                     cuda_graph=False):
    """Minibatch SVI training [Hensman et al. 2013] of the L-VAE (training.py:15-237).  Returns
    (penalty_term_arr, net_train_loss_arr, nll_loss_arr, recon_loss_arr, kld_loss_arr, m, H, best_epoch).
    cuda_graph=True (fixed T, natural-gradient mode): full minibatches run the GP side of the step as one CUDA-graph replay
    (graphed.GraphedHensmanStep); the last, shorter minibatch of an epoch takes the ordinary calls."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    from . import elbo_functions as _EF
    if _EF._GROUP is not None:
        # this loop is the reference's single-process loop: it passes the LOCAL N_batch // T as P_batch, while the sharded
        # bound needs the GLOBAL minibatch count (distributed.enable) — refuse rather than return a wrongly scaled bound
        raise RuntimeError("lvae_b200: hensman_training is a single-process loop; call distributed.disable() first, or drive "
                           "the sharded bound yourself (INTEGRATION.md, 'Multi-GPU': global P_batch + distributed.reduce_grads)")
    N = len(dataset)
    assert type_KL == 'GPapprox_closed'
    if varying_T:                                                                               # training.py:69-75
        n_batches = (P + subjects_per_batch - 1) // subjects_per_batch
        sampler = VaryingLengthBatchSampler(VaryingLengthSubjectSampler(dataset, id_covariate), subjects_per_batch)
    else:
        batch_size = subjects_per_batch * T
        n_batches = (P * T + batch_size - 1) // batch_size
        sampler = BatchSampler(SubjectSampler(dataset, P, T), batch_size, drop_last=False)
    dataloader = HensmanDataLoader(dataset, batch_sampler=sampler, num_workers=num_workers)

    gstep = None
    if cuda_graph:
        if varying_T or not natural_gradient:
            raise ValueError("lvae_b200: cuda_graph=True needs fixed T and natural_gradient=True")
        from .graphed import GraphedHensmanStep
        m = m.detach().to(torch.float64).contiguous().clone()              # updated in place from here on
        H = H.detach().to(torch.float64).contiguous().clone()
        gstep = GraphedHensmanStep(covar_module0, covar_module1, likelihoods, latent_dim, m, H, zt_list, P,
                                   subjects_per_batch, T, eps, natural_gradient_lr)

    curves = {k: [] for k in ("net", "recon", "nll", "kld", "penalty")}
    best_epoch = 0
    for epoch in range(1, epochs + 1):
        acc = None                                             # running [net, recon, nll, kld] / n_batches, kept on the device
        for sample_batched in dataloader:
            optimiser.zero_grad()
            nnet_model.train()
            covar_module0.train()
            covar_module1.train()
            data = sample_batched['digit'].double().to(device)
            train_x = sample_batched['label'].double().to(device)
            mask = sample_batched['mask'].double().to(device)
            N_batch = data.shape[0]

            recon_batch, mu, log_var = nnet_model(data)                                          # 103 (stock PyTorch VAE)
            recon_loss, nll = nnet_model.loss_function(recon_batch, data, mask)
            recon_loss, nll_loss = torch.sum(recon_loss), torch.sum(nll)
            PSD_H = H if natural_gradient else torch.matmul(H, H.transpose(-1, -2))              # 108
            graphed = gstep is not None and N_batch == subjects_per_batch * T
            if graphed:                                       # bound, all gradients and the update of (m, H): one replay
                P_in_current_batch = subjects_per_batch
                kld_loss = gstep(train_x, mu, log_var)
            elif varying_T:                                                                      # 110-115
                # the samplers deliver every subject as one contiguous block: count its rows on the loader's CPU batch
                # (no torch.unique on the device, no synchronisation); otherwise fall back to the reference's grouping
                counts = subject_counts_of(sample_batched['label'][:, id_covariate])
                P_in_current_batch = (len(counts) if counts is not None else
                                      torch.unique(train_x[:, id_covariate]).shape[0])
                kld_loss, grad_m, grad_H = minibatch_KLD_upper_bound_iter(
                    covar_module0, covar_module1, likelihoods, latent_dim, m, PSD_H, train_x, mu, log_var, zt_list, P,
                    P_in_current_batch, N, natural_gradient, id_covariate, eps, subject_counts=counts)
            else:
                P_in_current_batch = N_batch // T
                kld_loss, grad_m, grad_H = minibatch_KLD_upper_bound(
                    covar_module0, covar_module1, likelihoods, latent_dim, m, PSD_H, train_x, mu, log_var, zt_list, P,
                    P_in_current_batch, T, natural_gradient, eps)
            recon_loss = recon_loss * P / P_in_current_batch                                     # 117-124
            nll_loss = nll_loss * P / P_in_current_batch
            if loss_function == 'nll':
                net_loss = nll_loss + kld_loss
            elif loss_function == 'mse':
                kld_loss = kld_loss / latent_dim
                net_loss = recon_loss + weight * kld_loss
            net_loss.sum().backward()                                                            # 126-127
            optimiser.step()
            if natural_gradient and not graphed:                                                 # 129-135
                m2, H2 = natural_gradient_step(m, H, grad_m, grad_H, natural_gradient_lr)
                if gstep is not None:                         # keep the tensors the graph is bound to
                    m.copy_(m2.view_as(m))
                    H.copy_(H2)
                else:
                    m, H = m2, H2
            vals = torch.stack([net_loss.detach().sum(), recon_loss.detach(), nll_loss.detach(),
                                kld_loss.detach().sum()]).double() / n_batches                   # 137-140 without the four
            acc = vals if acc is None else acc + vals                                            # .item() syncs per step
        sums = dict(zip(("net", "recon", "nll", "kld"), acc.tolist()))                           # one sync per epoch
        if verbose:
            print('Iter %d/%d - Loss: %.3f  - GP loss: %.3f  - NLL Loss: %.3f  - Recon Loss: %.3f' % (
                epoch, epochs, sums["net"], sums["kld"], sums["nll"], sums["recon"]), flush=True)
        curves["penalty"].append(0.0)
        for k in ("net", "recon", "nll", "kld"):
            curves[k].append(sums[k])
        if (not epoch % 25) and epoch != epochs and on_validation is not None:                    # 150-235 (callback)
            if on_validation(epoch, dict(nnet_model=nnet_model, covar_module0=covar_module0, covar_module1=covar_module1,
                                         likelihoods=likelihoods, zt_list=zt_list, m=m, H=H)):
                best_epoch = epoch
    if gstep is not None:
        gstep.check_errors()
    if natural_gradient and device.type == "cuda":
        check_natural_gradient_errors()                       # the last steps' deferred failure flags
    arr = lambda k: np.asarray(curves[k], dtype=np.float64)
    return arr("penalty"), arr("net"), arr("nll"), arr("recon"), arr("kld"), m, H, best_epoch
