"""Host side of the Hensman minibatch step (training.py:108-135): the natural-gradient update of (m, H)."""
from . import ops


def natural_gradient_step(m, H, grad_m, grad_H, natural_gradient_lr, check=False):
    """training.py:129-135 in one launch: iH = H^-1; iH' = iH + lr (gH + gH^T); H <- iH'^-1;
    m <- H (iH m - lr (g_m - 2 gH m)).  Returns detached (m, H)."""
    m2, H2, info = ops.ng_step(m, H, grad_m, grad_H, natural_gradient_lr)
    if check and int(info[3].item()) != 0:
        raise RuntimeError(f"cholesky: natural-gradient update of latent {int(info[3].item()) - 1} is not positive-definite")
    return m2.detach(), H2.detach()
