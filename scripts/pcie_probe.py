"""How close are upload / download (cudaMemcpy2DAsync, host pitch != device pitch) to a plain contiguous PCIe copy?"""
import importlib, sys, time
import torch
sys.path.insert(0, ".")
cfd = importlib.import_module("macos-cfd_b200")
nx = ny = 8192
s = cfd.Solver(nx, ny, 1.0, 1.0)
sh = s.shapes()
hu, hv, hp, hr = (torch.zeros(sh[k], dtype=torch.float64).pin_memory() for k in ("u", "v", "p", "p"))
def t(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
up_b = (hu.numel() + hv.numel() + hp.numel()) * 8
dn_b = up_b + hr.numel() * 8
tu = t(lambda: s.upload(hu, hv, hp)); td = t(lambda: s.download(hu, hv, hp, hr))
print(f"upload   {tu*1e3:7.2f} ms  {up_b/tu/1e9:6.1f} GB/s"); print(f"download {td*1e3:7.2f} ms  {dn_b/td/1e9:6.1f} GB/s")
d = torch.empty(hu.numel(), dtype=torch.float64, device="cuda")
tu1 = t(lambda: d.copy_(hu.view(-1), non_blocking=True)); td1 = t(lambda: hu.view(-1).copy_(d, non_blocking=True))
print(f"contiguous H2D {hu.numel()*8/tu1/1e9:6.1f} GB/s   D2H {hu.numel()*8/td1/1e9:6.1f} GB/s")
# three solves at 8192^2 (BASELINE configs[4]: PCG vs MG pressure-solve time)
bc, Re, CFL = cfd.preset("jet", 1.0)
for _ in range(2): s.step(bc, Re, CFL, cfd.make_params("mg"))
import numpy as np
u, v, p, rhs = s.download()
stream = torch.cuda.ExternalStream(s.stream)
for name, pp in (("mg reference mode, 2 x smooth()", cfd.make_params("mg", vcycles=2, tol=0.0)),
                 ("PCG, 200 iterations", cfd.make_params("pcg", tol=0.0, pcg_max_iters=200)),
                 ("true V-cycle, 3 levels, 2 cycles, 2+2 sweeps", cfd.make_params("mg", vcycles=2, tol=0.0, mg_levels=3, rbgs_sweeps=2, mg_mode=1)),
                 ("true V-cycle, 8 levels, 2 cycles, 2+2 sweeps", cfd.make_params("mg", vcycles=2, tol=0.0, mg_levels=8, rbgs_sweeps=2, mg_mode=1))):
    ts = []
    for rep in range(3):
        s.upload(p=np.zeros_like(p), rhs=rhs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); res, it = s.pressure_solve(pp); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:48s} {min(ts):9.3f} ms  residual {res:.4e} iters {it}")
