"""Host-side helpers with the reference's names (utils/utils.py) plus batched versions.

* ``np_collate`` / ``keep_first_element`` — the reference's small utilities (:7-28), kept for scripts that import them.
* ``solve_continuous_are(A, B, Q, R, multiple_sol=False)`` — A^T P + P A - P B R^-1 B^T P + Q = 0 (:30-104).  The
  stabilising solution comes from the stable invariant subspace of the Hamiltonian matrix, computed with an ordered real
  Schur form like the reference does; ``multiple_sol=True`` enumerates the solutions spanned by other choices of n
  eigen-directions (the reference's debugging aid, :86-104).
* ``solve_continuous_are_batched`` — the stabilising solutions for MANY (A, B, Q, R) at once by the matrix sign function of
  the Hamiltonian (Newton iteration with determinant scaling, every step one batched inverse): torch, float64, on whichever
  device the inputs live.  Used to sweep cost weights / goals without a Python loop over SciPy calls.
* ``linearize_batched`` — central-difference Jacobians of a batched dynamics function about many (xf, uf) at once (what the
  reference obtains one point at a time with ``jax.jacobian``, controller/vhjb.py:156-160).
"""
from __future__ import annotations

from itertools import combinations
from typing import List, Union

import numpy as np


def np_collate(batch):
    """Collate NumPy samples (arrays, nested tuples / lists of arrays, scalars) into stacked arrays."""
    first = batch[0]
    if isinstance(first, np.ndarray):
        return np.stack(batch)
    if isinstance(first, (tuple, list)):
        return [np_collate(list(column)) for column in zip(*batch)]
    return np.array(batch)


def keep_first_element(func):
    """Decorator: a tuple result is reduced to its first element."""
    def wrapper(*args, **kwargs):
        out = func(*args, **kwargs)
        return out[0] if isinstance(out, tuple) else out
    return wrapper


def _hamiltonian(A, B, Q, R):
    S = B @ np.linalg.solve(R, B.T)
    return np.block([[A, -S], [-Q, -A.T]])


def solve_continuous_are(A, B, Q, R, multiple_sol: bool = False) -> Union[List[np.ndarray], np.ndarray]:
    """Computes in the inputs' own floating-point type, as the reference does (utils/utils.py:67-80): float32 inputs — what
    ``VHJBController.system_additional_init`` passes, controller/vhjb.py:159-160 — go through the single-precision Schur
    form; anything else is float64."""
    import scipy.linalg
    arrs = [np.asarray(v) for v in (A, B, Q, R)]
    dt = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64
    A, B, Q, R = (np.atleast_2d(a.astype(dt)) for a in arrs)
    n = A.shape[0]
    H = _hamiltonian(A, B, Q, R)
    if not multiple_sol:
        _, Z, sdim = scipy.linalg.schur(H, sort="lhp")          # stable eigenvalues first
        if sdim != n:
            raise np.linalg.LinAlgError("the Hamiltonian has eigenvalues on the imaginary axis: no stabilising solution")
        return Z[n:, :n] @ np.linalg.inv(Z[:n, :n])
    # every n-subset of eigenvectors whose top block is invertible spans a (generally indefinite / complex) solution;
    # the real symmetric ones are kept, rounded and de-duplicated like the reference's list
    w, V = np.linalg.eig(H)
    sols: List[np.ndarray] = []
    for idx in combinations(range(2 * n), n):
        X1, X2 = V[:n, idx], V[n:, idx]
        if abs(np.linalg.det(X1)) < 1e-10:
            continue
        P = X2 @ np.linalg.inv(X1)
        if np.abs(P.imag).max() > 1e-8:
            continue
        P = np.round(P.real, decimals=8) + 0.0
        if not any(np.array_equal(P, other) for other in sols):
            sols.append(P)
    return sols


def solve_continuous_are_batched(A, B, Q, R, iters: int = 60, tol: float = 1e-13):
    """Stabilising P [..., n, n] for batched A [..., n, n], B [..., n, m], Q [..., n, n], R [..., m, m] (broadcast over
    the leading dimensions).  Matrix sign function: Z <- (c Z + (c Z)^-1) / 2 with c = |det Z|^(-1/2n), Z0 = H; then
    [W12; W22 + I] P = -[W11 + I; W21] in the least-squares sense, W = sign(H)."""
    import torch
    f64 = torch.float64
    A, B, Q, R = (torch.as_tensor(v).to(f64) for v in (A, B, Q, R))
    dev = next((t.device for t in (A, B, Q, R) if t.is_cuda), A.device)
    A, B, Q, R = (t.to(dev) for t in (A, B, Q, R))
    n = A.shape[-1]
    lead = torch.broadcast_shapes(A.shape[:-2], B.shape[:-2], Q.shape[:-2], R.shape[:-2])
    A, B, Q, R = (t.expand(*lead, *t.shape[-2:]) for t in (A, B, Q, R))
    S = B @ torch.linalg.solve(R, B.transpose(-1, -2))
    Z = torch.cat([torch.cat([A, -S], dim=-1), torch.cat([-Q, -A.transpose(-1, -2)], dim=-1)], dim=-2)
    for _ in range(iters):
        c = torch.linalg.det(Z).abs().clamp_min(1e-300).pow(-1.0 / (2 * n))[..., None, None]
        Zs = c * Z
        Zn = 0.5 * (Zs + torch.linalg.inv(Zs))
        done = (Zn - Z).abs().amax(dim=(-1, -2)) <= tol * Zn.abs().amax(dim=(-1, -2))
        Z = Zn
        if bool(done.all()):
            break
    eye = torch.eye(n, dtype=f64, device=dev)
    lhs = torch.cat([Z[..., :n, n:], Z[..., n:, n:] + eye], dim=-2)
    rhs = -torch.cat([Z[..., :n, :n] + eye, Z[..., n:, :n]], dim=-2)
    P = torch.linalg.lstsq(lhs, rhs).solution
    return 0.5 * (P + P.transpose(-1, -2))


def lqr_gains_batched(A, B, Q, R):
    """(K [..., m, n], P [..., n, n]) with K = R^-1 B^T P for batched problems."""
    import torch
    P = solve_continuous_are_batched(A, B, Q, R)
    Bt = torch.as_tensor(B).to(P)
    Rt = torch.as_tensor(R).to(P)
    return torch.linalg.solve(Rt, Bt.transpose(-1, -2) @ P), P


def linearize_batched(xdot, xf, uf, eps: float = 1e-4):
    """Central-difference (A [N, n, n], B [N, n, m]) of ``xdot(x [K, n], u [K, m]) -> [K, n]`` about the rows of xf [N, n],
    uf [N, m]: all 2 (n + m) N perturbed evaluations go through ONE call (one ``hjb_dynamics`` launch when ``xdot`` is
    ``Dynamics.dynamics_step``)."""
    xf = np.atleast_2d(np.asarray(xf, dtype=np.float64))
    uf = np.atleast_2d(np.asarray(uf, dtype=np.float64))
    N, n = xf.shape
    m = uf.shape[1]
    d = n + m
    z = np.concatenate([xf, uf], axis=1)                                   # [N, d]
    pert = eps * np.eye(d)
    plus = (z[:, None, :] + pert[None]).reshape(N * d, d)
    minus = (z[:, None, :] - pert[None]).reshape(N * d, d)
    both = np.concatenate([plus, minus])
    f = np.asarray(xdot(both[:, :n], both[:, n:]), dtype=np.float64).reshape(2, N, d, n)
    J = ((f[0] - f[1]) / (2 * eps)).transpose(0, 2, 1)                     # [N, n, d]
    return J[:, :, :n], J[:, :, n:]
