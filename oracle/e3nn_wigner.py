"""Restatement of the e3nn functions the reference's SH rotation calls -- TEST INFRASTRUCTURE.

/root/reference/src/misc/sh_rotation.py:10-30 rotates the spherical-harmonic coefficients of every Gaussian with
``e3nn.o3.matrix_to_angles`` + ``e3nn.o3.wigner_D`` (e3nn is a pip dependency of the reference, requirements.txt; absent
from this image and not fetchable).  This file restates the published algorithm of e3nn 0.5.x ``o3/_wigner.py`` and
``o3/_rotation.py``: Y-X-Y Euler angles (the polar axis of e3nn's real harmonics is y), real-basis SO(3) generators
obtained from the su(2) ladder operators by ``change_basis_real_to_complex``, and
D^l(a, b, c) = exp(a X_y) exp(b X_x) exp(c X_y).

PARITY UNPINNED for this file: it cannot be checked against e3nn itself here.  What is checked (tests/test_adapter.py):
the representation property D(R1) D(R2) = D(R1 R2), orthogonality, D^1(R) = R (e3nn's l = 1 irrep is the vector (x, y, z)),
and the character  tr D^l(R) = sin((2l+1) w / 2) / sin(w / 2)  for rotation angle w.
"""
from __future__ import annotations

import math

import torch


def su2_generators(j: int) -> torch.Tensor:
    m = torch.arange(-j, j, dtype=torch.float64)
    raising = torch.diag(-torch.sqrt(j * (j + 1) - m * (m + 1)), diagonal=-1).to(torch.complex128)
    m = torch.arange(-j + 1, j + 1, dtype=torch.float64)
    lowering = torch.diag(torch.sqrt(j * (j + 1) - m * (m - 1)), diagonal=1).to(torch.complex128)
    m = torch.arange(-j, j + 1, dtype=torch.float64)
    return torch.stack([0.5 * (raising + lowering), torch.diag(1j * m.to(torch.complex128)), -0.5j * (raising - lowering)], dim=0)


def change_basis_real_to_complex(l: int) -> torch.Tensor:
    q = torch.zeros((2 * l + 1, 2 * l + 1), dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / 2 ** 0.5
        q[l + m, l - abs(m)] = -1j / 2 ** 0.5
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / 2 ** 0.5
        q[l + m, l - abs(m)] = 1j * (-1) ** m / 2 ** 0.5
    return (-1j) ** l * q


def so3_generators(l: int) -> torch.Tensor:
    X = su2_generators(l)
    Q = change_basis_real_to_complex(l)
    X = torch.conj(Q.T) @ X @ Q
    assert torch.all(torch.abs(torch.imag(X)) < 1e-9)
    return torch.real(X)


def matrix_x(a):
    c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
    return torch.stack([torch.stack([o, z, z], -1), torch.stack([z, c, -s], -1), torch.stack([z, s, c], -1)], -2)


def matrix_y(a):
    c, s, o, z = a.cos(), a.sin(), torch.ones_like(a), torch.zeros_like(a)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def angles_to_matrix(alpha, beta, gamma):
    alpha, beta, gamma = torch.broadcast_tensors(alpha, beta, gamma)
    return matrix_y(alpha) @ matrix_x(beta) @ matrix_y(gamma)


def xyz_to_angles(xyz):
    xyz = torch.nn.functional.normalize(xyz, p=2, dim=-1).clamp(-1, 1)
    return torch.atan2(xyz[..., 0], xyz[..., 2]), torch.acos(xyz[..., 1])


def matrix_to_angles(R):
    x = R @ R.new_tensor([0.0, 1.0, 0.0])
    a, b = xyz_to_angles(x)
    R = angles_to_matrix(a, b, torch.zeros_like(a)).transpose(-1, -2) @ R
    c = torch.atan2(R[..., 0, 2], R[..., 0, 0])
    return a, b, c


def wigner_D(l: int, alpha, beta, gamma):
    alpha, beta, gamma = torch.broadcast_tensors(alpha, beta, gamma)
    dt = alpha.dtype
    alpha = alpha[..., None, None] % (2 * math.pi)
    beta = beta[..., None, None] % (2 * math.pi)
    gamma = gamma[..., None, None] % (2 * math.pi)
    X = so3_generators(l).to(alpha.device)
    f = lambda ang, G: torch.matrix_exp(ang.double() * G)
    return (f(alpha, X[1]) @ f(beta, X[0]) @ f(gamma, X[1])).to(dt)


def rotate_sh(sh_coefficients, rotations):
    """/root/reference/src/misc/sh_rotation.py:10-30 on top of the restated e3nn functions."""
    n = sh_coefficients.shape[-1]
    alpha, beta, gamma = matrix_to_angles(rotations)
    out = []
    for degree in range(math.isqrt(n)):
        D = wigner_D(degree, alpha, beta, gamma).type(sh_coefficients.dtype)
        out.append(torch.einsum("...ij,...j->...i", D, sh_coefficients[..., degree ** 2:(degree + 1) ** 2]))
    return torch.cat(out, dim=-1)


def sh_rotation_blocks(rotations: torch.Tensor, degree: int) -> torch.Tensor:
    """[..., 3, 3] -> [..., (degree+1)^2, (degree+1)^2] block-diagonal matrix applied by rotate_sh."""
    alpha, beta, gamma = matrix_to_angles(rotations)
    n = (degree + 1) ** 2
    M = rotations.new_zeros((*rotations.shape[:-2], n, n))
    for l in range(degree + 1):
        M[..., l * l:(l + 1) ** 2, l * l:(l + 1) ** 2] = wigner_D(l, alpha, beta, gamma)
    return M
