// test_dropin.cpp -- exercises the reference-shaped C++ layer the way the reference's GUI loop does (main.cpp:200-264):
// eager coherence (the host State is current after every step), a host-side write between steps (what
// apply_shapes_to_simulation does, main.cpp:246), dt_override (gui.dt), and the public free functions.  It writes raw
// dumps that tests/test_gpu_cpp_driver.py compares with the oracle.   usage: test_dropin <outdir>
#include <cstdio>
#include <fstream>
#include <string>

#include "b200_backend.hpp"
#include "shapes.hpp"
#include "solver/time_integrator.hpp"
#include "viz.hpp"

static void dump(const std::string &path, const Field2D<double> &f) {
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char *>(f.data), static_cast<std::streamsize>(f.count() * sizeof(double)));
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    try {
        b200::set_coherence(b200::Coherence::Eager);
        Grid grid;
        grid.init(72, 40, 1.8, 1.0, 1);
        State st;
        st.u.allocate(grid.u_nx(), grid.ny, grid.u_pitch(), grid.ngx, grid.ngy);
        st.v.allocate(grid.nx, grid.v_ny(), grid.v_pitch(), grid.ngx, grid.ngy);
        st.p.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
        st.rhs.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
        PressureSolver pressure(grid);
        TimeIntegrator integ(grid);
        integ.clear_work();
        BC bc; // lid-driven cavity preset, main.cpp:288-296
        bc.left.type = bc.right.type = bc.bottom.type = BCType::Wall;
        bc.top.type = BCType::Moving;
        bc.top.moving = 1.0;
        PressureParams pp;
        pp.type = PressureSolverType::Multigrid;
        pp.vcycles = 2;
        double res = 0.0;
        std::ofstream log(dir + "/log.txt");
        log.precision(17);
        for (int step = 0; step < 6; ++step) {
            double dt = integ.step(st, bc, 1000.0, 0.1, pressure, pp, res, /*dt_override=*/0.0005);
            // the host State is current (eager): diagnostics and a GUI-style obstacle edit work on plain host arrays
            double d = divergence_l2(grid, st.u, st.v), m = max_velocity(grid, st.u, st.v);
            log << dt << ' ' << res << ' ' << d << ' ' << m << "\n";
            for (int jj = 10; jj < 14; ++jj) // "porous" obstacle: u *= 0.1 inside a small box (shapes.cpp:40-60)
                for (int ii = 20; ii < 26; ++ii) st.u.at_raw(ii, jj) *= 0.1;
        }
        dump(dir + "/u.bin", st.u), dump(dir + "/v.bin", st.v), dump(dir + "/p.bin", st.p), dump(dir + "/rhs.bin", st.rhs);

        // free functions on caller arrays
        Field2D<double> u2, v2, p2, du, rhs2;
        u2.allocate(grid.u_nx(), grid.ny, grid.u_pitch(), 1, 1);
        du.allocate(grid.u_nx(), grid.ny, grid.u_pitch(), 1, 1);
        v2.allocate(grid.nx, grid.v_ny(), grid.v_pitch(), 1, 1);
        p2.allocate(grid.nx, grid.ny, grid.p_pitch(), 1, 1);
        rhs2.allocate(grid.nx, grid.ny, grid.p_pitch(), 1, 1);
        for (size_t i = 0; i < u2.count(); ++i) u2.data[i] = st.u.data[i];
        for (size_t i = 0; i < v2.count(); ++i) v2.data[i] = st.v.data[i];
        for (size_t i = 0; i < p2.count(); ++i) p2.data[i] = st.p.data[i];
        BC jet;
        jet.left.type = BCType::Inflow, jet.right.type = BCType::Outflow;
        jet.jet_center = 0.5, jet.jet_width = 0.18;
        apply_bc_u(grid, u2, jet);
        apply_bc_v(grid, v2, jet);
        apply_bc_p(grid, p2, jet);
        log << compute_cfl(grid, u2, v2, 50.0, 0.01) << "\n";
        advect_u(grid, u2, v2, du);
        diffuse_u(grid, u2, du, 1.0 / 50.0);
        divergence(grid, u2, v2, rhs2);
        subtract_grad_p(grid, u2, v2, p2, 1e-3);
        PressureParams pcg;
        pcg.type = PressureSolverType::PCG;
        pcg.pcg_max_iters = 30;
        pcg.tol = 1e-9;
        PressureSolver solver2(grid);
        log << solver2.solve(p2, rhs2, pcg) << ' ' << solver2.last_iterations() << "\n";
        dump(dir + "/f_u.bin", u2), dump(dir + "/f_v.bin", v2), dump(dir + "/f_p.bin", p2), dump(dir + "/f_du.bin", du),
            dump(dir + "/f_rhs.bin", rhs2);

        // the GUI's per-frame readers / mutators (main.cpp:246, :255-264) with the reference's own signatures
        std::vector<Shape> shapes(2);
        shapes[0].type = 0, shapes[0].pos = {30.0f, 8.0f}, shapes[0].size = {9.0f, 5.0f}, shapes[0].bcType = 0;
        shapes[1].type = 1, shapes[1].pos = {50.5f, 20.25f}, shapes[1].size = {11.0f, 11.0f}, shapes[1].bcType = 1;
        apply_shape_boundary_conditions(grid, st, shapes);
        dump(dir + "/s_u.bin", st.u), dump(dir + "/s_v.bin", st.v);
        std::vector<double> speed;
        double vmin = 0.0, vmax = 0.0;
        compute_scalar(grid, st, ScalarField::Speed, speed, vmin, vmax);
        std::ofstream(dir + "/speed.bin", std::ios::binary)
            .write(reinterpret_cast<const char *>(speed.data()), static_cast<std::streamsize>(speed.size() * sizeof(double)));
        b200::FieldStats fs = b200::field_statistics(grid, st, ScalarField::Vorticity);
        log << vmin << ' ' << vmax << ' ' << fs.vmin << ' ' << fs.vmax << ' ' << fs.mean << ' ' << fs.stddev << "\n";
        std::vector<unsigned char> rgb;
        b200::render_rgb(grid, st, ScalarField::Speed, /*Turbo*/ 4, 0.0f, 1.0f, false, 1.0f, rgb);
        std::ofstream(dir + "/rgb.bin", std::ios::binary).write(reinterpret_cast<const char *>(rgb.data()), static_cast<std::streamsize>(rgb.size()));

        // lazy coherence: the State lives on the device.  A stand-alone PressureSolver::solve on other arrays (it shares the
        // Grid's device session) must not disturb it, and a new State that happens to get a freed State's addresses must
        // not be taken for the old one.
        b200::set_coherence(b200::Coherence::Lazy);
        for (int round = 0; round < 2; ++round) {
            State lz;
            lz.u.allocate(grid.u_nx(), grid.ny, grid.u_pitch(), grid.ngx, grid.ngy);
            lz.v.allocate(grid.nx, grid.v_ny(), grid.v_pitch(), grid.ngx, grid.ngy);
            lz.p.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
            lz.rhs.allocate(grid.nx, grid.ny, grid.p_pitch(), grid.ngx, grid.ngy);
            integ.clear_work();
            for (int step = 0; step < 6; ++step) {
                integ.step(lz, bc, 1000.0, 0.1, pressure, pp, res, 0.0005);
                if (step == 2) solver2.solve(p2, rhs2, pcg); // device p / rhs of the stepped State are ahead of the host here
            }
            b200::sync_to_host(lz);
            dump(dir + (round ? "/lazy2_u.bin" : "/lazy_u.bin"), lz.u), dump(dir + (round ? "/lazy2_p.bin" : "/lazy_p.bin"), lz.p);
        } // round 1 allocates its State where round 0's was freed
    } catch (const std::exception &e) {
        std::fprintf(stderr, "test_dropin: %s\n", e.what());
        return 1;
    }
    return 0;
}
