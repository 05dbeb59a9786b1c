import numpy as np
import torch


def maxdiff(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() if a.numel() else 0.0


def assert_close(a, b, atol, rtol=0.0, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel() == 0:
        return
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = (err > tol)
    assert not bad.any(), f"{what}: max|diff| {err.max().item():.3e} (tol {atol:g}+{rtol:g}*|ref|), {int(bad.sum())} bad"


def f_normalize(F):
    """Fundamental/essential matrices compare up to scale and sign: Frobenius-normalise, fix the sign on the
    largest-magnitude entry (SURVEY.md 8d parity thresholds)."""
    F = F.detach().double().cpu().reshape(-1, 9)
    F = F / F.norm(dim=1, keepdim=True)
    idx = F.abs().argmax(dim=1)
    sgn = torch.sign(F[torch.arange(F.shape[0]), idx])
    return (F * sgn[:, None]).reshape(-1, 3, 3)


def pose_set_distance(R1, R2, t, R1o, R2o, to):
    """{R1,R2} x {+-t} compared as a set: min over the two labelings of max Frobenius distance; t up to sign."""
    R1, R2, t, R1o, R2o, to = [x.detach().double().cpu() for x in (R1, R2, t, R1o, R2o, to)]
    t, to = t.reshape(-1, 3), to.reshape(-1, 3)
    a = torch.maximum((R1 - R1o).flatten(1).norm(dim=1), (R2 - R2o).flatten(1).norm(dim=1))
    b = torch.maximum((R1 - R2o).flatten(1).norm(dim=1), (R2 - R1o).flatten(1).norm(dim=1))
    dt = torch.minimum((t - to).norm(dim=1), (t + to).norm(dim=1))
    return torch.minimum(a, b), dt


def f_distance(A, B):
    """Sign/scale-invariant Frobenius distance between batches of 3x3 matrices."""
    A = A.detach().double().cpu().reshape(-1, 9)
    B = B.detach().double().cpu().reshape(-1, 9)
    A = A / A.norm(dim=1, keepdim=True)
    B = B / B.norm(dim=1, keepdim=True)
    return torch.minimum((A - B).norm(dim=1), (A + B).norm(dim=1))
