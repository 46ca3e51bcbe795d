"""Time one reference-mode mg solve (2 smooth() calls) for slab-shaped grids and forced strip heights."""
import importlib, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
cfd = importlib.import_module("macos-cfd_b200")
for nx, ny in ((1024, 1024), (1536, 1536), (2048, 1024), (512, 256), (1024, 512)):
    s = cfd.Solver(nx, ny, 1.0, 1.0 * ny / nx)
    rng = np.random.default_rng(1)
    sh = s.shapes()
    p0 = rng.uniform(-0.5, 0.5, sh["p"])
    s.upload(p=p0, rhs=rng.uniform(-0.5, 0.5, sh["p"]))
    stream = torch.cuda.ExternalStream(s.stream)
    pp = cfd.make_params("mg", vcycles=2, tol=0.0)
    out = []
    for strip in (0, 17, 33, 49, 65, 97, -2):
        if strip == -2:
            s.set_variant(2); s.set_smoother_strip(0)
        else:
            s.set_variant(1); s.set_smoother_strip(strip)
        ts = []
        for rep in range(6):
            s.upload(p=p0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.pressure_solve(pp)
            e1.record(stream)
            e1.synchronize()
            if rep >= 2:
                ts.append(e0.elapsed_time(e1))
        out.append((strip, round(min(ts) / 2 * 1e3, 1), s.last_smoother[7:]))
    print(nx, ny, out, flush=True)
    s.close()
