#!/usr/bin/env python
"""Generates tests/golden/oracle_vectors.npz: small seeded input -> output vectors of the CPU oracle (oracle/cts_ref.py +
oracle/cts_models.py, itself pinned to the reference's known answers in reference_kats.json) for the BASELINE configs in
miniature.  The GPU parity tests compare the CUDA path with these files, so a drift of the oracle and a drift of the
kernels are caught independently.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cts_models as M  # noqa: E402
import cts_ref as R  # noqa: E402

# (key, algorithm name, Newton kwargs, constraint, nonlinear, halo, dtype, steps, dt, shape (nlev, nx, ny))
COLUMN_CASES = [
    ("c3_ars343_f64", "ARS343", dict(max_iters=1, update_j="NewTimeStep"), None, False, 1, np.float64, 3, 0.02, (64, 6, 5)),
    ("c3_ars343_f32", "ARS343", dict(max_iters=1, update_j="NewTimeStep"), None, False, 1, np.float32, 3, 0.02, (64, 6, 5)),
    ("c3_ars222_f64", "ARS222", dict(max_iters=1), None, False, 0, np.float64, 2, 0.02, (32, 5, 4)),
    ("ssp333_ssprk_f64", "SSP333", dict(max_iters=2), "SSP", True, 1, np.float64, 2, 0.01, (16, 4, 4)),
    ("ssp333_ssprk_f32", "SSP333", dict(max_iters=2), "SSP", True, 1, np.float32, 2, 0.01, (16, 4, 4)),
]


def column_case(name, newton, constraint, nonlinear, halo, dtype, steps, dt, shape):
    nlev, nx, ny = shape
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, halo, dtype)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=0.7, halo=halo, nonlinear=nonlinear)
    kw = dict(newton)
    if "update_j" in kw:
        kw["update_j"] = R.UpdateEvery(kw["update_j"])
    alg = R.imex_algorithm(name, R.NewtonsMethod(**kw), constraint)
    u = T0.copy()
    model.dss(u, None, 0.0)
    _, step = R.make_stepper(model.ode_function(), alg, u)
    t = 0.0
    for _ in range(steps):
        step(u, None, t, dt)
        t += dt
    return u


def main():
    out = {}
    for key, name, newton, constraint, nonlinear, halo, dtype, steps, dt, shape in COLUMN_CASES:
        out[key] = column_case(name, newton, constraint, nonlinear, halo, dtype, steps, dt, shape)
    # C2 in miniature: LSRK54CarpenterKennedy on periodic upwind advection
    n = 4096
    u = M.fill_hash(n, 1)
    du = np.zeros(n)
    adv = M.Advection(n)
    t = 0.0
    for _ in range(3):
        R.lsrk_step(u, du, None, t, 0.5 / n, adv.incr, R.lsrk54_tableau())
        t += 0.5 / n
    out["c2_lsrk54_f64"] = u
    # Rosenbrock (SSPKnoth) on the linear column problem
    nlev, nx, ny = 32, 6, 5
    K, T0, S = M.column_inputs(nlev, nx, ny, 0, ny, 1, np.float64)
    model = M.ColumnModel(nlev, nx, ny, K, S_lev=S, nu=0.7, halo=1)
    f = R.ClimaODEFunction(T_exp=model.T_exp, T_imp=R.ODEFunction(model.T_imp, jac_prototype=model.jac_prototype().zero(), Wfact=model.Wfact),
                           dss=model.dss)
    u = T0.copy()
    _, step = R.make_stepper(f, R.RosenbrockAlgorithm(R.ssp_knoth()), u)
    t = 0.0
    for _ in range(3):
        step(u, None, t, 0.01)
        t += 0.01
    out["rosenbrock_sspknoth_f64"] = u
    # batched tridiagonal solve (K4): W = dtγJ - I of the column model, hashed right-hand side
    Wm = model.jac_prototype().zero()
    model.Wfact(Wm, T0, None, 0.0087, 0.0)
    rhs = M.fill_hash(model.n, 9)
    x = np.zeros(model.n)
    Wm.ldiv(x, rhs)
    out["k4_tridiag_x_f64"] = x
    path = os.path.join(HERE, "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.dtype.name, v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
