// state.hpp -- the caller-owned State (reference: src/solver/state.hpp:5-9)
#pragma once
#include "field.hpp"

struct State {
    Field2D<double> u, v, p;  // primary fields
    Field2D<double> rhs, tmp; // work buffers (tmp is never touched by the solver)
    Field2D<float> scalar;    // visualisation buffer (never touched by the solver)
};
