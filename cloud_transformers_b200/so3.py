"""so3_exponential_map in plain torch (Rodrigues' formula, squared angle clamped to >= eps) -- what the reference
gets from pytorch3d.transforms.so3 (layers/utils.py:6,29,56).  Twelve floats per head; stays in PyTorch so autograd
carries d R / d log_R, the kernels only consume R."""
import torch


def so3_exponential_map(log_rot, eps: float = 0.0001):
    nrms = (log_rot * log_rot).sum(1)
    theta = torch.clamp(nrms, eps).sqrt()
    fac1 = theta.sin() / theta
    fac2 = (1.0 - theta.cos()) / (theta * theta)
    x, y, z = log_rot.unbind(1)
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]
    return fac1[:, None, None] * K + fac2[:, None, None] * torch.bmm(K, K) + eye
