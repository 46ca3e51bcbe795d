"""One rank of a row-slab run (launched by torchrun from tests/test_gpu_multi.py or by hand):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/slab_worker.py \
        --preset jet --solver mg --nx 256 --ny 192 --steps 4 --out /tmp/slab.npz

Every rank steps its slab; rank 0 gathers the owned rows, assembles the global u, v, p, rhs and writes them (plus the
per-step dt / residual / iteration history) to --out.  The caller compares them with a single-GPU run and the oracle.
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def viz_shapes(cfd, nx, ny):
    """A solid box in the lower part, a porous circle that straddles the middle of the grid (slab boundaries), one disabled."""
    return [cfd.Shape(0, nx * 0.2, ny * 0.1, nx * 0.1, ny * 0.15, 0, 1), cfd.Shape(1, nx * 0.5, ny * 0.42, ny * 0.2, ny * 0.2, 1, 1),
            cfd.Shape(0, 1.0, 1.0, 5.0, 5.0, 0, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="jet")
    ap.add_argument("--solver", default="mg")
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=192)
    ap.add_argument("--Lx", type=float, default=2.0)
    ap.add_argument("--Ly", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--vcycles", type=int, default=2)
    ap.add_argument("--variant", type=int, default=1)
    ap.add_argument("--strip", type=int, default=0)  # > 0: force the streaming smoother with this strip height
    ap.add_argument("--p2p", type=int, default=1)  # 0: force the NCCL path (CFD_B200_P2P=0)
    ap.add_argument("--viz", type=int, default=0)  # 1: obstacle masks + visualisation scalars / coloured frame after the steps
    ap.add_argument("--bt-periodic", type=int, default=0)  # 1: bottom and top sides Periodic instead of the preset's
    ap.add_argument("--mg-mode", type=int, default=0)  # 1: the true V-cycle extension
    ap.add_argument("--mg-levels", type=int, default=3)
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--delay-rank", type=int, default=-1)  # this rank's stream is kept busy before every step (test hook)
    ap.add_argument("--delay-us", type=float, default=0.0)
    ap.add_argument("--reset-after", type=int, default=-1)  # collective cfd_b200_comm_reset after that many steps
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    os.environ["CFD_B200_P2P"] = str(a.p2p)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")  # CPU plumbing only: the id broadcast and the final gather
    cfd = importlib.import_module("macos-cfd_b200")
    ids = [cfd.comm_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s = cfd.Solver(a.nx, a.ny, a.Lx, a.Ly, device=local % torch.cuda.device_count(), rank=rank, nranks=world,
                   comm_id=ids[0])
    s.set_variant(a.variant)
    s.set_smoother_strip(a.strip)
    bc, Re, CFL = cfd.preset(a.preset, a.Ly)
    if a.bt_periodic:
        bc.bottom.type = bc.top.type = cfd.PERIODIC
    pp = cfd.make_params(a.solver, vcycles=a.vcycles, tol=a.tol, pcg_max_iters=a.iters, mg_levels=a.mg_levels,
                         rbgs_sweeps=a.sweeps, mg_mode=a.mg_mode)
    r0, n = s.row0, s.nrows
    if a.preset == "shear":
        s.upload(u=np.ascontiguousarray(cfd.shear_initial_u(a.nx, a.ny, a.Ly)[r0:r0 + n + 2]))
    hist = []
    for k in range(a.steps):
        if k == a.reset_after:
            s.comm_reset()
        if rank == a.delay_rank % world and a.delay_us > 0:
            s.debug_delay(a.delay_us)  # the neighbours run ahead as far as the flag protocol lets them
        info = s.step(bc, Re, CFL, pp)
        hist.append((info.dt, info.residual, info.iters, info.div_before, info.div_after))
    extra = {}
    if a.viz:  # the GUI's per-frame mutators / readers on slabs: shapes in cells of the whole grid, whole-grid statistics
        s.apply_shapes(viz_shapes(cfd, a.nx, a.ny))
        for name in ("speed", "vorticity", "pressure"):
            out, lo, hi, mean, std = s.compute_scalar(cfd.FIELDS[name])
            extra[name], extra[name + "_stats"] = out, np.array([lo, hi, mean, std])
        extra["rgb"] = s.render_rgb(cfd.FIELDS["speed"], 4)
        info = s.step(bc, Re, CFL, pp)  # and a step on the masked State
    div_l2, vmax = s.divergence_l2(), s.max_velocity()
    u, v, p, rhs = s.download()
    # owned rows: u, p, rhs local rows 1..n; v rows 1..n (+ n+1 on the top slab); ghost rows from the end slabs
    top, bottom = rank == world - 1, rank == 0
    part = {"u": u[(0 if bottom else 1):(n + 2 if top else n + 1)], "v": v[(0 if bottom else 1):(n + 3 if top else n + 1)],
            "p": p[(0 if bottom else 1):(n + 2 if top else n + 1)], "rhs": rhs[(0 if bottom else 1):(n + 2 if top else n + 1)]}
    for k in ("speed", "vorticity", "pressure", "rgb"):
        if k in extra:
            part[k] = extra[k]
    parts = [None] * world
    dist.gather_object(part, parts if rank == 0 else None, dst=0)
    if rank == 0:
        out = {k: np.concatenate([q[k] for q in parts], axis=0) for k in part}
        out.update({k: v for k, v in extra.items() if k.endswith("_stats")})
        np.savez(a.out, hist=np.array(hist), div_l2=div_l2, vmax=vmax, comm_mode=s.comm_mode, **out)
    s.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
