// viz.hpp -- the reference's visualisation reader (src/viz.hpp:8-16) with the same names and signatures, answered by the
// device (cfd_b200_compute_scalar).  When `s` is the State a TimeIntegrator is stepping in lazy-coherence mode, nothing
// crosses PCIe except the nx*ny result; otherwise u, v, p are uploaded first.
#pragma once
#include <vector>

#include "solver/grid.hpp"
#include "solver/state.hpp"

enum class ScalarField { U, V, Speed, Pressure, Vorticity, Streamlines, Temperature }; // viz.hpp:8

// viz.hpp:11-12 / viz.cpp:6-65
void compute_scalar(const Grid &g, const State &s, ScalarField field, std::vector<double> &out, double &vmin, double &vmax);

namespace b200 {
// Gui::update_field_statistics (gui.cpp:1433-1481): min / max / mean / standard deviation without the field itself.
struct FieldStats {
    double vmin = 0, vmax = 0, mean = 0, stddev = 0;
};
FieldStats field_statistics(const Grid &g, const State &s, ScalarField field);
// update_field_texture (gui.cpp:1505-1561) without the GL upload: nx*ny*3 bytes, colormap = Gui::Colormap as int.
void render_rgb(const Grid &g, const State &s, ScalarField field, int colormap, float scale_min, float scale_max,
                bool normalize, float gamma, std::vector<unsigned char> &rgb);
} // namespace b200
