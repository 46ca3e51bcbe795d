// shapes.hpp -- the reference's obstacle masks (src/shapes.hpp:13, src/shapes.cpp:7-74) on the device
// (cfd_b200_apply_shapes).  The reference takes `Shape` from gui/gui.hpp (ImGui types); the GUI is out of scope here,
// so the struct is restated with plain floats: same members, same meaning (gui.hpp:14-23).
#pragma once
#include <vector>

#include "solver/grid.hpp"
#include "solver/state.hpp"

struct Shape {
    int type = 0;                 // 0 = rectangle, 1 = circle
    struct Vec2 { float x = 0.0f, y = 0.0f; } pos, size{10.0f, 10.0f};
    int bcType = 0;               // 0 = wall, 1 = porous
    bool visible = true;
    bool enabled = true;
};

void apply_shape_boundary_conditions(const Grid &grid, State &state, const std::vector<Shape> &shapes);
