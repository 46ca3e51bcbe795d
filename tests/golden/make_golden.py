#!/usr/bin/env python
"""Generate tests/golden/*.npz|json from oracle/_ref (the reference's own sources compiled here).

Run in the dev container (needs /root/reference):   OMP_NUM_THREADS=1 python tests/golden/make_golden.py
The outputs are committed; the GPU box (no /root/reference) only reads them.

Files
  runs.json    per (preset, solver, grid): per-step (dt, divergence_l2, residual), final norms, sha256(u|v|p)
               -- the 64x32 entries reproduce the known answers of SURVEY.md §8c
  fields.npz   final u, v, p, rhs of every run in runs.json
  stages.npz   seeded random inputs and the reference's outputs for every stage function
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["OMP_NUM_THREADS"] = "1"
from oracle import cfdo  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (tag, nx, ny, Lx, Ly, steps, tol, max_iters, vcycles)
GRIDS = [
    ("64x32", 64, 32, 2.0, 1.0, 5, 1e-8, 200, 2),      # SURVEY §8c golden grid (power-of-two spacing)
    ("40x24", 40, 24, 1.3, 0.7, 3, 1e-8, 60, 2),       # non-power-of-two spacing, dx != dy
]


def run_case(lib, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, vcycles):
    bc, Re, CFL = cfdo.preset(preset, Ly)
    pp = cfdo.make_params(solver, vcycles=vcycles, tol=tol, pcg_max_iters=iters)
    sim = cfdo.Sim(lib, nx, ny, Lx, Ly)
    if preset == "shear":
        cfdo.shear_init(sim.u, nx, ny, Ly)
    hist = []
    for _ in range(steps):
        dt = sim.step(bc, Re, CFL, pp)
        div = lib.divergence_l2(nx, ny, Lx, Ly, sim.u, sim.v)
        hist.append([dt, div, sim.last_residual])
    out = {k: getattr(sim, k).copy() for k in ("u", "v", "p", "rhs")}
    sim.close()
    return hist, out


def main():
    lib = cfdo.load("ref")
    assert lib.L.cfdo_kind() == b"reference"
    runs, fields = {}, {}
    for tag, nx, ny, Lx, Ly, steps, tol, iters, vcycles in GRIDS:
        for preset in ("jet", "lid", "shear"):
            for solver in ("pcg", "mg"):
                hist, out = run_case(lib, preset, solver, nx, ny, Lx, Ly, steps, tol, iters, vcycles)
                key = f"{preset}_{solver}_{tag}"
                sha = hashlib.sha256(out["u"].tobytes() + out["v"].tobytes() + out["p"].tobytes()).hexdigest()
                runs[key] = dict(preset=preset, solver=solver, nx=nx, ny=ny, Lx=Lx, Ly=Ly, steps=steps, tol=tol,
                                 pcg_max_iters=iters, vcycles=vcycles, history=hist, sha256_uvp=sha,
                                 norm_u=float(np.linalg.norm(out["u"])), norm_v=float(np.linalg.norm(out["v"])),
                                 norm_p=float(np.linalg.norm(out["p"])))
                for k, a in out.items():
                    fields[f"{key}/{k}"] = a
                print(key, sha[:16], "res %.6e" % hist[-1][2])
    with open(os.path.join(HERE, "runs.json"), "w") as f:
        json.dump(runs, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **fields)

    # ---- stage-level vectors on a small odd-sized grid with random data
    nx, ny, Lx, Ly = 23, 17, 1.7, 0.9
    sh = cfdo.shapes(nx, ny)
    rng = np.random.default_rng(20261019)
    st = {"grid": np.array([nx, ny, Lx, Ly])}
    u = rng.uniform(-0.5, 0.5, sh["u"])
    v = rng.uniform(-0.5, 0.5, sh["v"])
    p = rng.uniform(-0.5, 0.5, sh["p"])
    st["u"], st["v"], st["p"] = u, v, p
    du = np.zeros(sh["u"]); dv = np.zeros(sh["v"])
    lib.advect_u(nx, ny, Lx, Ly, u, v, du); lib.advect_v(nx, ny, Lx, Ly, u, v, dv)
    st["advect_u"], st["advect_v"] = du.copy(), dv.copy()
    lib.diffuse_u(nx, ny, Lx, Ly, u, du, 1.0 / 50.0); lib.diffuse_v(nx, ny, Lx, Ly, v, dv, 1.0 / 50.0)
    st["rhs_u"], st["rhs_v"] = du, dv
    st["cfl"] = np.array([lib.compute_cfl(nx, ny, Lx, Ly, u, v, 50.0, 0.5),
                          lib.compute_cfl(nx, ny, Lx, Ly, u * 1e-3, v * 1e-3, 1000.0, 0.1)])
    rhs = np.zeros(sh["p"]); lib.divergence(nx, ny, Lx, Ly, u, v, rhs); st["divergence"] = rhs
    u2, v2 = u.copy(), v.copy(); lib.subtract_grad_p(nx, ny, Lx, Ly, u2, v2, p, 1e-3)
    st["gradp_u"], st["gradp_v"] = u2, v2
    st["div_l2"] = np.array([lib.divergence_l2(nx, ny, Lx, Ly, u, v)])
    st["max_vel"] = np.array([lib.max_velocity(nx, ny, Lx, Ly, u, v)])
    for solver in ("pcg", "mg"):
        ps = rng.uniform(-0.5, 0.5, sh["p"])  # warm-start contents (mg keeps them, pcg clears them)
        st[f"solve_{solver}_p0"] = ps.copy()
        res, _ = lib.pressure_solve(nx, ny, Lx, Ly, ps, rhs, cfdo.make_params(solver, vcycles=3, tol=1e-9,
                                                                              pcg_max_iters=25))
        st[f"solve_{solver}_p"], st[f"solve_{solver}_res"] = ps, np.array([res])
    combos = [(cfdo.INFLOW, cfdo.OUTFLOW, cfdo.WALL, cfdo.WALL), (cfdo.WALL, cfdo.WALL, cfdo.WALL, cfdo.MOVING),
              (cfdo.PERIODIC, cfdo.PERIODIC, cfdo.WALL, cfdo.WALL), (cfdo.OUTFLOW, cfdo.INFLOW, cfdo.PERIODIC,
                                                                    cfdo.PERIODIC),
              (cfdo.MOVING, cfdo.PERIODIC, cfdo.OUTFLOW, cfdo.INFLOW)]
    st["bc_combos"] = np.array(combos)
    for n, c in enumerate(combos):
        bc = cfdo.make_bc(*c, left_inflow_u=0.8, jet_center=0.45, jet_width=0.3, left_moving=0.11,
                          right_moving=-0.12, bottom_moving=0.13, top_moving=1.0, right_inflow_u=0.21,
                          right_inflow_v=0.22, bottom_inflow_u=0.23, bottom_inflow_v=0.24, top_inflow_u=0.25,
                          top_inflow_v=0.26)
        for name, a in (("u", u), ("v", v), ("p", p)):
            b = a.copy(); lib.apply_bc(name, nx, ny, Lx, Ly, b, bc); st[f"bc{n}_{name}"] = b
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **st)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
