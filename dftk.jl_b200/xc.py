"""Exchange-correlation energy densities and potentials on the device (the XC dispatch point of
ext/DFTKCUDAExt.jl:17-25; formulas: Dirac exchange, VWN5, PW92, PBE with libxc's constants).

Evaluated with torch float64 elementwise kernels; the potentials are exact derivatives of the energy
expression obtained with torch.autograd (so e, Vρ and Vσ are always mutually consistent).  This is SCF
plumbing adjacent to the hot path, not one of the graded kernels.
"""
import math
import torch

DENS_THRESHOLD = 1e-15
_RS = (3 / (4 * math.pi)) ** (1 / 3)
_FZ_DEN = 2 ** (4 / 3) - 2
_FPP0 = 4 / (9 * (2 ** (1 / 3) - 1))


def _cbrt(x):
    return torch.pow(x, 1.0 / 3.0)


def _fzeta(z):
    return ((1 + z) ** (4 / 3) + (1 - z) ** (4 / 3) - 2) / _FZ_DEN


def _ex_unif(n):
    return -(3 / 4) * (3 / math.pi) ** (1 / 3) * n * _cbrt(n)


_VWN = ((0.0310907, 3.72744, 12.9352, -0.10498), (0.01554535, 7.06042, 18.0578, -0.32500),
        (-1 / (6 * math.pi ** 2), 1.13107, 13.0045, -0.0047584))


def _vwn_piece(x, A, b, c, x0):
    Q = math.sqrt(4 * c - b * b)
    X = x * x + b * x + c
    X0 = x0 * x0 + b * x0 + c
    at = torch.atan(Q / (2 * x + b))
    return A * (torch.log(x * x / X) + (2 * b / Q) * at
                - (b * x0 / X0) * (torch.log((x - x0) ** 2 / X) + (2 * (b + 2 * x0) / Q) * at))


def _ec_vwn(rs, zeta):
    x = torch.sqrt(rs)
    p = [_vwn_piece(x, *par) for par in (_VWN if zeta is not None else _VWN[:1])]
    if zeta is None:
        return p[0]
    fz, z4 = _fzeta(zeta), zeta ** 4
    return p[0] + p[2] * fz * (1 - z4) / _FPP0 + (p[1] - p[0]) * fz * z4


_PW_A = {"pw": ((0.0310907, 0.01554535, 0.0168869), 1.709921),
         "pw_mod": ((0.0310906908696548950, 0.01554534543482744750, 0.0168868639404617),
                    1.709920934161365617563962776245)}
_PW_C = ((0.21370, 7.5957, 3.5876, 1.6382, 0.49294), (0.20548, 14.1189, 6.1977, 3.3662, 0.62517),
         (0.11125, 10.357, 3.6231, 0.88026, 0.49671))


def _pw_G(rs, i, a):
    a1, b1, b2, b3, b4 = _PW_C[i]
    s = torch.sqrt(rs)
    den = 2 * a * (b1 * s + b2 * rs + b3 * rs * s + b4 * rs * rs)
    return -2 * a * (1 + a1 * rs) * torch.log(1 + 1 / den)


def _ec_pw(rs, zeta, kind):
    a, fz20 = _PW_A[kind]
    g0 = _pw_G(rs, 0, a[0])
    if zeta is None:
        return g0
    g1, mac = _pw_G(rs, 1, a[1]), _pw_G(rs, 2, a[2])
    fz, z4 = _fzeta(zeta), zeta ** 4
    return g0 - mac * fz * (1 - z4) / fz20 + (g1 - g0) * fz * z4


_KAPPA, _BETA = 0.8040, 0.06672455060314922
_MU = _BETA * (math.pi ** 2 / 3)
_GAMMA = (1 - math.log(2)) / math.pi ** 2


def _ex_pbe(n, sigma):
    kF = _cbrt(3 * math.pi ** 2 * n)
    s2 = sigma / (4 * kF * kF * n * n)
    return _ex_unif(n) * (1 + _KAPPA - _KAPPA / (1 + _MU * s2 / _KAPPA))


def _ec_pbe(n, rs, zeta, sigma):
    ec = _ec_pw(rs, zeta, "pw_mod")
    if zeta is None:
        phi2, phi3 = 1.0, 1.0
    else:
        phi = ((1 + zeta) ** (2 / 3) + (1 - zeta) ** (2 / 3)) / 2
        phi2, phi3 = phi * phi, phi ** 3
    kF = _cbrt(3 * math.pi ** 2 * n)
    t2 = sigma / (4 * phi2 * (4 * kF / math.pi) * n * n)
    A = (_BETA / _GAMMA) / (torch.exp(-ec / (_GAMMA * phi3)) - 1)
    At2 = A * t2
    return ec + _GAMMA * phi3 * torch.log(1 + (_BETA / _GAMMA) * t2 * (1 + At2) / (1 + At2 + At2 * At2))


def evaluate(functionals, rho, sigma=None):
    """rho: (n_spin, N); sigma: (1|3, N) or None.  Returns e (N,), Vrho (n_spin,N), Vsigma or None."""
    n_spin = rho.shape[0]
    is_gga = any(f.startswith("gga") for f in functionals)
    mask = rho.sum(dim=0) > DENS_THRESHOLD
    r = torch.where(mask, rho, torch.full_like(rho, 1.0 / n_spin)).detach().requires_grad_(True)
    sg = None
    if is_gga:
        sg = torch.where(mask, sigma, torch.zeros_like(sigma)).detach().requires_grad_(True)
    if n_spin == 1:
        n, zeta = r[0], None
    else:
        n = r[0] + r[1]
        zeta = torch.clamp((r[0] - r[1]) / n, -1 + 1e-14, 1 - 1e-14)
    rs = _RS / _cbrt(n)
    e = torch.zeros_like(n)
    for f in functionals:
        if f == "lda_x":
            e = e + (_ex_unif(n) if n_spin == 1 else
                     0.5 * (_ex_unif(2 * r[0].clamp_min(1e-30)) + _ex_unif(2 * r[1].clamp_min(1e-30))))
        elif f == "lda_c_vwn":
            e = e + n * _ec_vwn(rs, zeta)
        elif f == "lda_c_pw":
            e = e + n * _ec_pw(rs, zeta, "pw")
        elif f == "gga_x_pbe":
            e = e + (_ex_pbe(n, sg[0]) if n_spin == 1 else
                     0.5 * (_ex_pbe(2 * r[0].clamp_min(1e-30), 4 * sg[0]) + _ex_pbe(2 * r[1].clamp_min(1e-30), 4 * sg[2])))
        elif f == "gga_c_pbe":
            e = e + n * _ec_pbe(n, rs, zeta, sg[0] if n_spin == 1 else sg[0] + 2 * sg[1] + sg[2])
        else:
            raise NotImplementedError(f"functional {f}")
    grads = torch.autograd.grad(e.sum(), [r] + ([sg] if is_gga else []))
    z = torch.zeros((), dtype=rho.dtype, device=rho.device)
    ev = torch.where(mask, e.detach(), z)
    vr = torch.where(mask, grads[0], z)
    vs = torch.where(mask, grads[1], z) if is_gga else None
    return ev, vr, vs
