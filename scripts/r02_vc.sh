cd /root/repo; mkdir -p gpurun_out
python - <<PY
import importlib, sys, time, numpy as np, torch
sys.path.insert(0, ".")
cfd = importlib.import_module("macos-cfd_b200")
nx = ny = 8192
s = cfd.Solver(nx, ny, 1.0, 1.0)
bc, Re, CFL = cfd.preset("jet", 1.0)
for _ in range(2): s.step(bc, Re, CFL, cfd.make_params("mg"))
u, v, p, rhs = s.download()
stream = torch.cuda.ExternalStream(s.stream)
for name, pp in (("V-cycle 8 levels 2x V(2,2)", cfd.make_params("mg", vcycles=2, tol=0.0, mg_levels=8, rbgs_sweeps=2, mg_mode=1)),
                 ("V-cycle 3 levels", cfd.make_params("mg", vcycles=2, tol=0.0, mg_levels=3, rbgs_sweeps=2, mg_mode=1))):
    ts = []
    for rep in range(4):
        s.upload(p=np.zeros_like(p), rhs=rhs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); res, it = s.pressure_solve(pp); e1.record(stream); e1.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    print(name, ts, res, it)
PY
python -m pytest tests/test_vcycle.py -m gpu -q 2>&1 | tail -2
