// pressure.hpp -- PressureSolver with the reference's interface (reference: src/solver/pressure.hpp:6-38).
#pragma once
#include "field.hpp"
#include "grid.hpp"

enum class PressureSolverType { Multigrid, PCG };

struct PressureParams {
    PressureSolverType type = PressureSolverType::Multigrid;
    int mg_levels = 3;   // only read by the V-cycle extension (the reference never reads it)
    int vcycles = 2;
    int rbgs_sweeps = 2; // only read by the V-cycle extension
    double tol = 1e-6;
    int pcg_max_iters = 200;
    // B200 extension, default = reference behaviour: false -> what the reference's "Multigrid" really is (fine-grid
    // red-black Gauss-Seidel, pressure.cpp:113-129); true -> a true geometric V-cycle (parity unpinned).
    bool true_vcycle = false;
};

namespace b200 { struct Session; }

class PressureSolver {
  public:
    explicit PressureSolver(const Grid &g);
    ~PressureSolver();
    PressureSolver(const PressureSolver &) = delete;
    PressureSolver &operator=(const PressureSolver &) = delete;

    // Solve Laplace(p) = rhs on the device; returns the final residual norm (1e9 when non-finite).
    double solve(Field2D<double> &p, const Field2D<double> &rhs, const PressureParams &params);
    double last_residual() const { return last_residual_; }
    int last_iterations() const { return last_iters_; } // extension: the reference does not expose its count

  private:
    friend struct TimeIntegrator;
    const Grid &g;
    b200::Session *session_;
    double last_residual_ = 0.0;
    int last_iters_ = 0;
    bool warned_nan_ = false;
};
