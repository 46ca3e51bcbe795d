// time_integrator.hpp -- the RK2 + projection integrator with the reference's interface (reference:
// src/solver/time_integrator.hpp:13-30).  step() runs entirely on the device; see b200_backend.hpp for how the
// caller-owned host State stays coherent.
#pragma once
#include "advection.hpp"
#include "bc.hpp"
#include "diffusion.hpp"
#include "grid.hpp"
#include "metrics.hpp"
#include "pressure.hpp"
#include "project.hpp"
#include "state.hpp"

// Replace non-finite values by 0 (reference: time_integrator.cpp:8-27).  Inside step() this is folded into the
// producing kernels; the free function is a host utility for callers.
void sanitize_field(Field2D<double> &field);

struct TimeIntegrator {
    const Grid &g;
    // The reference materialises these work arrays; the B200 path fuses the derivatives into the Runge-Kutta kernels and
    // keeps u1/v1 on the device.  The members exist (allocated, zero) for source compatibility only.
    Field2D<double> du_dt, dv_dt;
    Field2D<double> u1, v1;

    explicit TimeIntegrator(const Grid &grid);
    ~TimeIntegrator();
    void clear_work();

    // Advance one step.  Returns the chosen dt; pressure_residual is an output.
    double step(State &s, const BC &bc, double Re, double CFL, PressureSolver &pressure, const PressureParams &pp,
                double &pressure_residual, double dt_override = 0.0);

  private:
    b200::Session *session_;
};
