// cfd2d_b200 -- the headless driver of the reference (`cfd2d_gui --no-gui`, src/main.cpp:52-159) on the B200 path.
//
// Same flags, same defaults, same stdout line every 10 steps, same out/final.vtk.  The reference hard-codes the grid
// (512x256 on 2x1), the Jet Plume preset and 2000 steps (main.cpp:105, :131-142, :147); those become flags here:
//   --nx= --ny= --Lx= --Ly= --preset=jet|lid|shear --steps= --vtk=<path|none> --dump=<path> --print-every=
//   --mg-mode=reference|vcycle  --device=
// The loop itself is the reference's: integrator.step(...), divergence_l2(...), printf.  In the default (lazy)
// coherence mode the State stays on the device and nothing crosses PCIe inside the loop.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>

#include "b200_backend.hpp"
#include "solver/time_integrator.hpp"

// Legacy ASCII VTK writer, byte-compatible with the reference's save_vtk (main.cpp:23-50): STRUCTURED_POINTS,
// cell-centred pressure scalars and velocity vectors, default ostream formatting.
static void write_vtk(const Grid &g, const State &s, const std::string &path) {
    const auto dir = std::filesystem::path(path).parent_path();
    if (!dir.empty()) std::filesystem::create_directories(dir);
    std::ofstream out(path);
    out << "# vtk DataFile Version 3.0\nCFD2D snapshot\nASCII\n"
        << "DATASET STRUCTURED_POINTS\n"
        << "DIMENSIONS " << g.nx << " " << g.ny << " 1\n"
        << "ORIGIN 0 0 0\n"
        << "SPACING " << g.dx << " " << g.dy << " 1\n"
        << "POINT_DATA " << g.nx * g.ny << "\n"
        << "SCALARS pressure double\nLOOKUP_TABLE default\n";
    for (int j = 0; j < g.ny; ++j)
        for (int i = 0; i < g.nx; ++i) out << s.p.at_raw(i + g.ngx, j + g.ngy) << "\n";
    out << "VECTORS velocity double\n";
    for (int j = 0; j < g.ny; ++j)
        for (int i = 0; i < g.nx; ++i) {
            const int ii = i + g.ngx, jj = j + g.ngy;
            const double uc = 0.5 * (s.u.at_raw(ii, jj) + s.u.at_raw(ii + 1, jj));
            const double vc = 0.5 * (s.v.at_raw(ii, jj) + s.v.at_raw(ii, jj + 1));
            out << uc << ' ' << vc << " 0\n";
        }
}

// raw dump for parity tooling: u | v | p | rhs in the reference host layout
static void write_dump(const State &s, const std::string &path) {
    std::ofstream out(path, std::ios::binary);
    for (const Field2D<double> *f : {&s.u, &s.v, &s.p, &s.rhs})
        out.write(reinterpret_cast<const char *>(f->data), static_cast<std::streamsize>(f->count() * sizeof(double)));
}

static bool flag(const std::string &arg, const char *name, std::string &val) {
    const size_t n = std::strlen(name);
    if (arg.compare(0, n, name) != 0) return false;
    val = arg.substr(n);
    return true;
}

int main(int argc, char **argv) {
    // defaults of the reference's main(), main.cpp:54-59
    double Re = 50.0, CFL_target = 0.01;
    PressureParams pp;
    pp.type = PressureSolverType::PCG;
    pp.pcg_max_iters = 200;
    pp.tol = 1e-6;
    int nx = 512, ny = 256, steps = 2000, print_every = 10;
    double Lx = 2.0, Ly = 1.0;
    std::string preset = "jet", vtk = "out/final.vtk", dump;
    bool Re_set = false, CFL_set = false;

    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        std::string v;
        if (arg == "--no-gui") { // accepted: this driver is always headless
        } else if (flag(arg, "--Re=", v)) Re = std::stod(v), Re_set = true;
        else if (flag(arg, "--CFL=", v)) CFL_target = std::stod(v), CFL_set = true;
        else if (flag(arg, "--solver=", v)) pp.type = v == "pcg" ? PressureSolverType::PCG : PressureSolverType::Multigrid;
        else if (flag(arg, "--mg-levels=", v)) pp.mg_levels = std::stoi(v);
        else if (flag(arg, "--vcycles=", v)) pp.vcycles = std::stoi(v);
        else if (flag(arg, "--tol=", v)) pp.tol = std::stod(v);
        else if (flag(arg, "--pcg-iters=", v)) pp.pcg_max_iters = std::stoi(v);
        else if (arg == "--tvd-debug") tvd_debug = true, std::fprintf(stderr, "TVD debugging enabled\n");
        else if (flag(arg, "--nx=", v)) nx = std::stoi(v);
        else if (flag(arg, "--ny=", v)) ny = std::stoi(v);
        else if (flag(arg, "--Lx=", v)) Lx = std::stod(v);
        else if (flag(arg, "--Ly=", v)) Ly = std::stod(v);
        else if (flag(arg, "--preset=", v)) preset = v;
        else if (flag(arg, "--steps=", v)) steps = std::stoi(v);
        else if (flag(arg, "--print-every=", v)) print_every = std::stoi(v);
        else if (flag(arg, "--vtk=", v)) vtk = v;
        else if (flag(arg, "--dump=", v)) dump = v;
        else if (flag(arg, "--mg-mode=", v)) pp.true_vcycle = (v == "vcycle");
        else if (flag(arg, "--device=", v)) b200::set_device(std::stoi(v));
        else if (arg == "--help") {
            std::printf("cfd2d_b200 - headless 2D incompressible solver on NVIDIA B200\n"
                        "  --Re= --CFL= --solver=pcg|mg --mg-levels= --vcycles= --tol= --pcg-iters= --tvd-debug   (reference flags)\n"
                        "  --nx= --ny= --Lx= --Ly= --preset=jet|lid|shear --steps= --print-every= --vtk=<path|none> --dump=<path>\n"
                        "  --mg-mode=reference|vcycle --device=\n");
            return 0;
        } else {
            std::fprintf(stderr, "unknown option %s\n", arg.c_str());
            return 2;
        }
    }

    try {
        b200::set_coherence(b200::Coherence::Lazy);
        Grid grid;
        grid.init(nx, ny, Lx, Ly, 1);
        State state; // main.cpp:106-112
        state.u.allocate(grid.u_nx(), grid.ny, grid.u_pitch(), grid.ngx, grid.ngy);
        state.v.allocate(grid.nx, grid.v_ny(), grid.v_pitch(), grid.ngx, grid.ngy);
        state.p.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
        state.rhs.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
        state.tmp.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
        state.scalar.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);

        PressureSolver pressure(grid);
        TimeIntegrator integrator(grid);
        integrator.clear_work();

        BC bc; // presets: main.cpp:131-142 (jet), :288-296 (lid), :297-304 + :348-358 (shear)
        if (preset == "jet") {
            bc.left.type = BCType::Inflow, bc.right.type = BCType::Outflow;
            bc.top.type = bc.bottom.type = BCType::Wall;
            bc.left.inflow_u = 1.0, bc.left.inflow_v = 0.0;
            bc.jet_center = 0.5 * grid.Ly, bc.jet_width = 0.18 * grid.Ly, bc.jet_thickness = 0.02 * grid.Ly;
        } else if (preset == "lid") {
            bc.left.type = bc.right.type = bc.bottom.type = BCType::Wall;
            bc.top.type = BCType::Moving, bc.top.moving = 1.0;
            if (!Re_set) Re = 1000.0;
            if (!CFL_set) CFL_target = 0.1;
        } else if (preset == "shear") {
            bc.left.type = bc.right.type = BCType::Periodic;
            bc.bottom.type = bc.top.type = BCType::Wall;
            if (!Re_set) Re = 1000.0;
            if (!CFL_set) CFL_target = 0.1;
            for (int j = 0; j < grid.ny; ++j) {
                const double y = (j + 0.5) * grid.dy;
                const double val = bc.left.inflow_u * std::sin(2.0 * 3.141592653589793 * y / grid.Ly);
                for (int i = 0; i < grid.u_nx(); ++i) state.u.at_raw(i + grid.ngx, j + grid.ngy) = val;
            }
        } else {
            std::fprintf(stderr, "unknown preset %s\n", preset.c_str());
            return 2;
        }

        double sim_time = 0.0, pres_res = 0.0;
        for (int step = 0; step < steps; ++step) { // main.cpp:147-156
            double dt = integrator.step(state, bc, Re, CFL_target, pressure, pp, pres_res);
            sim_time += dt;
            double divL2 = divergence_l2(grid, state.u, state.v);
            if (print_every > 0 && step % print_every == 0)
                std::printf("step %d time %g dt %g div %g res %g\n", step, sim_time, dt, divL2, pres_res);
        }
        b200::sync_to_host(state);
        if (vtk != "none") write_vtk(grid, state, vtk);
        if (!dump.empty()) write_dump(state, dump);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "cfd2d_b200: %s\n", e.what());
        return 1;
    }
    return 0;
}
