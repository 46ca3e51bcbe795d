// metrics.hpp (reference: src/solver/metrics.hpp:6-13)
#pragma once
#include "field.hpp"
#include "grid.hpp"

double compute_cfl(const Grid &, const Field2D<double> &u, const Field2D<double> &v, double Re, double CFL_target);
double max_velocity(const Grid &, const Field2D<double> &u, const Field2D<double> &v);
double divergence_l2(const Grid &, const Field2D<double> &u, const Field2D<double> &v);
