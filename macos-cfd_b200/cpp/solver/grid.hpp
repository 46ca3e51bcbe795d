// grid.hpp -- MAC-grid geometry with the reference's public surface (reference: src/solver/grid.hpp:3-24,
// grid.cpp:3-12).  Plain host-side bookkeeping; the device layout lives behind the C ABI.
#pragma once

struct Grid {
    int nx = 0, ny = 0;        // cells
    int ngx = 1, ngy = 1;      // ghost layers (the B200 path supports the reference's only value, 1)
    double Lx = 1.0, Ly = 1.0; // domain size
    double dx = 0.0, dy = 0.0; // cell size

    void init(int Nx, int Ny, double Lx_, double Ly_, int ng = 1) {
        nx = Nx, ny = Ny, Lx = Lx_, Ly = Ly_, ngx = ngy = ng;
        dx = Lx / static_cast<double>(nx);
        dy = Ly / static_cast<double>(ny);
    }

    // cell centres
    int p_pitch() const { return nx + 2 * ngx; }
    int p_idx(int i, int j) const { return (j + ngy) * p_pitch() + (i + ngx); }
    // x-faces
    int u_nx() const { return nx + 1; }
    int u_pitch() const { return u_nx() + 2 * ngx; }
    int u_idx(int i, int j) const { return (j + ngy) * u_pitch() + (i + ngx); }
    // y-faces
    int v_ny() const { return ny + 1; }
    int v_pitch() const { return nx + 2 * ngx; }
    int v_idx(int i, int j) const { return (j + ngy) * v_pitch() + (i + ngx); }
};
