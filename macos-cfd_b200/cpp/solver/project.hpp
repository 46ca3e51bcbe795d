// project.hpp (reference: src/solver/project.hpp:8-12)
#pragma once
#include "field.hpp"
#include "grid.hpp"

void divergence(const Grid &g, const Field2D<double> &u, const Field2D<double> &v, Field2D<double> &rhs);
void subtract_grad_p(const Grid &g, Field2D<double> &u, Field2D<double> &v, const Field2D<double> &p, double dt);
