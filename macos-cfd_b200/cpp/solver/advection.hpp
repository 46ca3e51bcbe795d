// advection.hpp (reference: src/solver/advection.hpp:7-10)
#pragma once
#include "field.hpp"
#include "grid.hpp"

// Accepted for source compatibility.  The reference uses it to log clip overshoots (advection.cpp:19-27); on the
// B200 path the clip is provably inactive for finite data, so there is nothing to log.
extern bool tvd_debug;

void advect_u(const Grid &, const Field2D<double> &u, const Field2D<double> &v, Field2D<double> &du_dt);
void advect_v(const Grid &, const Field2D<double> &u, const Field2D<double> &v, Field2D<double> &dv_dt);
