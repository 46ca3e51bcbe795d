// bc.hpp -- boundary-condition description and the three apply_bc_* entry points with the reference's signatures
// (reference: src/solver/bc.hpp:6-34).  The functions run the B200 kernels k_bc_lr / k_bc_bt through the C ABI.
#pragma once
#include "field.hpp"
#include "grid.hpp"

enum class BCType { Wall, Moving, Inflow, Outflow, Periodic };

struct BCSide {
    BCType type = BCType::Wall;
    double moving = 0.0;   // tangential speed for Moving
    double inflow_u = 0.0; // imposed inflow velocity
    double inflow_v = 0.0;
};

struct BC {
    BCSide left{BCType::Wall, 0.0, 1.0, 0.0};
    BCSide right, bottom, top;
    // jet inflow on the left side; the reference only ever reads jet_center and jet_width (bc.cpp:7-29)
    double jet_center = 0.5;
    double jet_width = 0.2;
    double jet_thickness = 0.02;
    double jet_eps = 0.03;
    int jet_k = 8;
    double jet_phase = 0.0;
};

double jet_velocity(const Grid &g, const BC &bc, double y);
void apply_bc_u(const Grid &, Field2D<double> &, const BC &);
void apply_bc_v(const Grid &, Field2D<double> &, const BC &);
void apply_bc_p(const Grid &, Field2D<double> &, const BC &);
