"""Host side of the Hensman minibatch step (training.py:108-135): the natural-gradient update of (m, H)."""
from . import ops


def natural_gradient_step(m, H, grad_m, grad_H, natural_gradient_lr, check=False):
    """training.py:129-135 in one launch: iH = H^-1; iH' = iH + lr (gH + gH^T); H <- iH'^-1;
    m <- H (iH m - lr (g_m - 2 gH m)).  Returns detached (m, H)."""
    hinv = None
    tag = getattr(grad_H, "_lvae_hinv", None)          # H^-1 computed by the bound for this very H (same storage, unmodified)
    if tag is not None and tag[1] == H.data_ptr() and tag[2] == H._version and H.dtype == tag[0].dtype:
        hinv = tag[0]
    m2, H2, info = ops.ng_step(m, H, grad_m, grad_H, natural_gradient_lr, Hinv=hinv)
    if check and int(info[3].item()) != 0:
        raise RuntimeError(f"cholesky: natural-gradient update of latent {int(info[3].item()) - 1} is not positive-definite")
    return m2.detach(), H2.detach()
