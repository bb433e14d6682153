"""numpy restatements of the reference's own test fixtures (``/root/reference/test``), used to pin
the oracle.  Each builder cites the fixture it follows.  Test infrastructure only."""
from __future__ import annotations

import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import cts_ref as R  # noqa: E402


# ------------------------------------------------------------------ single column (hydrostatic) ---
class SingleColumn:
    """test/integration/single_column_ars.jl:6-146, 154-209 (N=30, L=30e3, dt=100)."""

    def __init__(self, N=30, L=30e3):
        self.N, self.L = N, L
        self.MSLP, self.grav, self.R_m, self.gamma = 1e5, 9.80616, 287.0024093890231, 1.4
        self.C_p = self.R_m * self.gamma / (self.gamma - 1)
        self.C_v = self.R_m / (self.gamma - 1)
        self.dz = np.zeros(N + 1) + L / N
        rho = np.zeros(N)
        rhotheta = np.zeros(N)
        for i in range(1, N + 1):
            z = (i - 0.5) * L / N
            Tv, p, r = self._profile(z, 300.0, 300.0)
            rho[i - 1] = r
            rhotheta[i - 1] = r * Tv * (self.MSLP / p) ** (self.R_m / self.C_p)
        self.u0 = np.concatenate([rho, rhotheta, np.zeros(N + 1)])

    def P(self, rt):  # :168
        return self.MSLP ** (-self.R_m / self.C_v) * self.R_m ** (self.C_p / self.C_v) * rt ** (self.C_p / self.C_v)

    def dP(self, rt):  # :169-170
        return (self.C_p / self.C_v * self.MSLP ** (-self.R_m / self.C_v) * self.R_m ** (self.C_p / self.C_v)
                * rt ** (self.C_p / self.C_v - 1))

    def _profile(self, z, T_virt_surf, T_min_ref):  # :43-60
        H_sfc = self.R_m * T_virt_surf / self.grav
        H_t = H_sfc
        zp = z / H_t
        th = math.tanh(zp)
        dTv = T_virt_surf - T_min_ref
        Tv = T_virt_surf - dTv * th
        dTvp = dTv / T_virt_surf
        p = -H_t * (zp + dTvp * (math.log(1 - dTvp * th) - math.log(1 + th) + zp))
        p /= H_sfc * (1 - dTvp ** 2)
        p = self.MSLP * math.exp(p)
        return Tv, p, p / (self.R_m * Tv)

    def residual(self, du, u, p, t):  # spatial_residual! :62-92
        N = self.N
        dz = self.L / N
        rho, rt, w = u[:N], u[N:2 * N], u[2 * N:3 * N + 1]
        rho_res, rt_res, w_res = np.zeros(N), np.zeros(N), np.zeros(N + 1)
        for i in range(N):
            rp = 0.0 if i == N - 1 else (rho[i] + rho[i + 1]) / 2.0
            rm = 0.0 if i == 0 else (rho[i] + rho[i - 1]) / 2.0
            rho_res[i] = -(w[i + 1] * rp - w[i] * rm) / dz
            tp = 0.0 if i == N - 1 else (rt[i] + rt[i + 1]) / 2.0
            tm = 0.0 if i == 0 else (rt[i] + rt[i - 1]) / 2.0
            rt_res[i] = -(w[i + 1] * tp - w[i] * tm) / dz
        for i in range(1, N):
            w_res[i] = -(self.P(rt[i]) - self.P(rt[i - 1])) / (1 / 2 * (rho[i - 1] + rho[i])) / dz - self.grav
        du[:N], du[N:2 * N], du[2 * N:] = rho_res, rt_res, w_res

    def jacobian(self, J, u):  # jacobian! :94-135 (0-based here)
        N = self.N
        dz = self.dz
        rho, rt = u[:N], u[N:2 * N]
        rh = np.concatenate([[rho[0]], (rho[:-1] + rho[1:]) / 2.0, [rho[-1]]])
        th = np.concatenate([[rt[0]], (rt[:-1] + rt[1:]) / 2.0, [rt[-1]]])
        pc = np.array([self.P(x) for x in rt])
        dpc = np.array([self.dP(x) for x in rt])
        dzh = np.concatenate([[np.nan], (dz[:N - 1] + dz[1:N]) / 2.0, [np.nan]])
        for i in range(N):
            J[i, i + 2 * N] = rh[i] / dz[i]
            J[i, i + 2 * N + 1] = -rh[i + 1] / dz[i]
            J[i + N, i + 2 * N] = th[i] / dz[i]
            J[i + N, i + 2 * N + 1] = -th[i + 1] / dz[i]
        for i in range(1, N):
            J[i + 2 * N, i - 1] = (pc[i] - pc[i - 1]) / dzh[i] / (2 * rh[i] ** 2)
            J[i + 2 * N, i] = (pc[i] - pc[i - 1]) / dzh[i] / (2 * rh[i] ** 2)
            J[i + 2 * N, i - 1 + N] = dpc[i - 1] / (rh[i] * dzh[i])
            J[i + 2 * N, i + N] = -dpc[i] / (rh[i] * dzh[i])

    def wfact(self, W, u, p, dtg, t):  # single_column_Wfact! :137-146
        W.M[...] = 0.0
        self.jacobian(W.M, u)
        W.M *= dtg
        W.M[np.diag_indices_from(W.M)] -= 1.0

    def ode_function(self):  # :319-324
        n = 3 * self.N + 1

        def T_exp(du, u, p, t):
            du[...] = 0.0

        T_imp = R.ODEFunction(self.residual, jac_prototype=R.DenseJacobian(n), Wfact=self.wfact)
        return R.ClimaODEFunction(T_exp=T_exp, T_imp=T_imp)

    def in_test_ars343_reference(self, dt=100.0):
        """The independent ARS343 comparator inside the reference test (:234-297), one step."""
        N = self.N
        n = 3 * N + 1
        g = 0.4358665215084590
        a42 = a43 = 0.5529291480359398
        b1 = -3 / 2 * g ** 2 + 4 * g - 1 / 4
        b2 = 3 / 2 * g ** 2 - 5 * g + 5 / 4
        a31 = ((1 - 9 / 2 * g + 3 / 2 * g ** 2) * a42 + (11 / 4 - 21 / 2 * g + 15 / 4 * g ** 2) * a43
               - 7 / 2 + 13 * g - 9 / 2 * g ** 2)
        a32 = ((-1 + 9 / 2 * g - 3 / 2 * g ** 2) * a42 + (-11 / 4 + 21 / 2 * g - 15 / 4 * g ** 2) * a43
               + 4 - 25 / 2 * g + 9 / 2 * g ** 2)
        a41 = 1 - a42 - a43
        a_exp = np.array([[0, 0, 0, 0], [g, 0, 0, 0], [a31, a32, 0, 0], [a41, a42, a43, 0]])
        b_exp = np.array([0, b1, b2, g])
        a_imp = np.array([[0, 0, 0, 0], [0, g, 0, 0], [0, (1 - g) / 2, g, 0], [0, b1, b2, g]])
        b_imp = b_exp.copy()
        ns = 3
        K_exp, K_imp = np.zeros((n, ns + 1)), np.zeros((n, ns + 1))
        u = self.u0.copy()

        def f_imp_solve(u0, gam):
            k0 = np.zeros(n)
            J0 = np.zeros((n, n))
            self.residual(k0, u0, None, 0.0)
            self.jacobian(J0, u0)
            return np.linalg.solve(np.eye(n) - gam * J0, k0)

        u_stage = u.copy()
        for st in range(1, ns + 1):
            K_exp[:, st - 1] = 0.0
            u_tmp = u + dt * (K_imp[:, :st - 1] @ a_imp[st, 1:st] + K_exp[:, :st] @ a_exp[st, :st])
            K_imp[:, st - 1] = f_imp_solve(u_tmp, dt * a_imp[st, st])
            u_stage = u_tmp + dt * K_imp[:, st - 1] * a_imp[st, st]
        u += dt * (K_exp[:, :ns + 1] @ b_exp[:ns + 1] + K_imp[:, :ns] @ b_imp[1:ns + 1])
        return float(np.linalg.norm(u))


# ------------------------------------------------------------------ ARKode analytic problems ---
class IntegratorTestCase:
    """test/utils/test_case_utils.jl:8-62 (ClimaIntegratorTestCase): builds the un-split and split problems."""

    def __init__(self, name, linear_implicit, t_end, Y0, analytic_sol, tendency, Wfact_dense,
                 implicit_tendency=None, explicit_tendency=None, dtype=np.float64):
        self.name, self.linear_implicit, self.t_end = name, linear_implicit, t_end
        self.Y0 = np.array(Y0, dtype=dtype)
        self.analytic_sol = analytic_sol
        n = len(Y0)

        def wf(W, Y, p, dtg, t):
            W.M[...] = Wfact_dense(Y, dtg, t)

        mk = lambda f: R.ODEFunction(f, jac_prototype=R.DenseJacobian(n, dtype), Wfact=wf)
        self.f = R.ClimaODEFunction(T_imp=mk(tendency))
        self.split_f = R.ClimaODEFunction(
            T_exp=explicit_tendency, T_imp=mk(tendency if implicit_tendency is None else implicit_tendency))


def ark_analytic(dtype=np.float64):
    """test/problems/arkode_references.jl:11-27: y' = λy + 1/(1+t²) − λ atan(t), y = atan(t)."""
    lam = -100.0
    src = lambda t: 1 / (1 + t ** 2) - lam * math.atan(t)

    def tend(Yt, Y, p, t):
        Yt[...] = lam * Y + src(t)

    def imp(Yt, Y, p, t):
        Yt[...] = lam * Y

    def exp_(Yt, Y, p, t):
        Yt[...] = src(t)

    return IntegratorTestCase("ark_analytic", True, 10.0, [0.0], lambda t: np.array([math.atan(t)]), tend,
                              lambda Y, dtg, t: np.array([[dtg * lam - 1]]), imp, exp_, dtype)


def ark_analytic_nonlin(dtype=np.float64):
    """arkode_references.jl:35-47: y' = (t+1) exp(-y)."""

    def tend(Yt, Y, p, t):
        Yt[...] = (t + 1) * np.exp(-Y)

    return IntegratorTestCase("ark_analytic_nonlin", False, 10.0, [0.0],
                              lambda t: np.array([math.log(t ** 2 / 2 + t + 1)]), tend,
                              lambda Y, dtg, t: np.array([[(-dtg * (t + 1) * math.exp(-Y[0]) - 1)]]), dtype=dtype)


def ark_analytic_sys(dtype=np.float64):
    """arkode_references.jl:52-69: Y' = A Y, A = V D V^-1, D = diag(-1/2, -1/10, -100) (BASELINE config 1)."""
    V = np.array([[1, -1, 1], [-1, 2, 1], [0, -1, 2]], dtype=dtype)
    Vi = np.array([[5, 1, -3], [2, 2, -2], [1, 1, 1]], dtype=dtype) / 4
    D = np.diag(np.array([-1 / 2, -1 / 10, -100.0], dtype=dtype))
    A = V @ D @ Vi
    Y0 = np.array([1, 1, 1], dtype=dtype)

    def tend(Yt, Y, p, t):
        Yt[...] = A @ Y

    sol = lambda t: V @ np.diag(np.exp(np.diag(D) * t)) @ Vi @ Y0
    return IntegratorTestCase("ark_analytic_sys", True, 1 / 20, Y0, sol, tend,
                              lambda Y, dtg, t: dtg * A - np.eye(3, dtype=dtype), dtype=dtype)


def stiff_linear(dtype=np.float64):
    """test/problems/stiff_linear_ode.jl:18-36."""
    lam = 1000.0
    A_imp = np.array([[-lam, 0], [0, 0]], dtype=dtype)
    A_exp = np.array([[0, 1], [1, -1]], dtype=dtype)
    A = A_imp + A_exp
    Y0 = np.array([1, 1], dtype=dtype)
    w, Q = np.linalg.eig(A.astype(np.float64))

    def sol(t):
        return np.real(Q @ np.diag(np.exp(w * t)) @ np.linalg.solve(Q, Y0.astype(np.float64)))

    return IntegratorTestCase(
        "stiff_linear", True, 0.01, Y0, sol, lambda Yt, Y, p, t: Yt.__setitem__(Ellipsis, A @ Y),
        lambda Y, dtg, t: dtg * A_imp - np.eye(2, dtype=dtype),
        lambda Yt, Y, p, t: Yt.__setitem__(Ellipsis, A_imp @ Y),
        lambda Yt, Y, p, t: Yt.__setitem__(Ellipsis, A_exp @ Y), dtype)


# ------------------------------------------------------------------ LSRK scalar problems ---
def linear_prob(dtype=np.float64):
    """test/problems/linear_odes.jl:9-23: du/dt = αu, u0 = 1/2, α = 1.01."""
    alpha_p = 1.01

    def f(du, u, p, t, a=True, b=False):
        du[...] = a * p * u + b * du

    return f, np.array([0.5], dtype=dtype), (0.0, 1.0), alpha_p, lambda u0, p, t: u0 * math.exp(p * t)


def sincos_prob():
    """linear_odes.jl:29-44."""

    def f(du, u, p, t, a=True, b=False):
        d0 = a * p * u[1] + b * du[0]
        d1 = -a * p * u[0] + b * du[1]
        du[0], du[1] = d0, d1

    def sol(u0, p, t):
        s, c = math.sin(p * t), math.cos(p * t)
        return np.array([[c, s], [-s, c]]) @ u0

    return f, np.array([0.0, 1.0]), (0.0, 1.0), 2.0, sol
