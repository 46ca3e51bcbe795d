// kernels_viz.cuh -- the host-side readers / mutators around the time step, on the device:
//   k_viz_scalar   compute_scalar (viz.cpp:6-65) + the sums of Gui::update_field_statistics (gui.cpp:1452-1474)
//   k_viz_rgb      the pixel loop of update_field_texture (gui.cpp:1523-1551) with Colormaps::apply_colormap
//                  (colormaps.cpp:114-127) -- a frame can be coloured without the fields ever crossing PCIe
//   k_shape_apply  apply_shape_boundary_conditions (shapes.cpp:7-74): obstacle masks on u, v after a step
// All three are HBM-trivial next to a step (one pass over 2-3 fields); they exist so that an interactive caller never
// has to download the State per frame.  Arithmetic follows the reference expression by expression (fp64 for the scalar,
// fp32 for the colour and the circle test, no FMA contraction), so the scalar field, the masks and -- for gamma = 1 --
// the bytes are bit-identical; with gamma != 1 std::pow (glibc) and the device pow may differ in the last place.
#pragma once
#include "common.cuh"

enum { VIZ_U = 0, VIZ_V, VIZ_SPEED, VIZ_PRESSURE, VIZ_VORTICITY, VIZ_STREAMLINES, VIZ_TEMPERATURE }; // viz.hpp:8

// one cell of compute_scalar, viz.cpp:20-58; o = element offset of the cell in the device layout
__device__ __forceinline__ double viz_value(const Geom &g, const double *__restrict__ u, const double *__restrict__ v,
                                            const double *__restrict__ p, int field, size_t o) {
    double val = 0.0;
    switch (field) {
    case VIZ_U: val = 0.5 * (u[o] + u[o + 1]); break;
    case VIZ_V: val = 0.5 * (v[o] + v[o + g.P]); break;
    case VIZ_SPEED: {
        const double uc = 0.5 * (u[o] + u[o + 1]), vc = 0.5 * (v[o] + v[o + g.P]);
        val = sqrt(uc * uc + vc * vc);
        break;
    }
    case VIZ_PRESSURE:
    case VIZ_TEMPERATURE: val = p[o]; break; // Temperature is a placeholder for pressure (viz.cpp:50-53)
    case VIZ_VORTICITY: {
        const double dv_dx = (v[o + 1] - v[o - 1]) / (2.0 * g.dx);
        const double du_dy = (u[o + g.P] - u[o - g.P]) / (2.0 * g.dy);
        val = dv_dx - du_dy;
        break;
    }
    default: val = 0.0; // Streamlines: placeholder (viz.cpp:45-49)
    }
    return fin0(val);
}

// out[i + j*nx] = the scalar; partial minima / maxima / sums per block, finished by the last block in block order:
// sc->viz = {min, max, sum, sum of squares} (min / max start from +-1e30 like viz.cpp:11-12).
__global__ void __launch_bounds__(256)
    k_viz_scalar(Geom g, const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ p,
                 int field, double *__restrict__ out, Scalars *sc, double *partials) { pdl_begin();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double lo = 1e30, hi = -1e30, acc[2] = {0.0, 0.0};
    if (i < g.nx) {
        for (int j = blockIdx.y; j < g.ny; j += gridDim.y) {
            const double val = viz_value(g, u, v, p, field, gidx(g, i + 1, j + 1));
            out[(size_t)j * g.nx + i] = val;
            lo = min2(lo, val);
            hi = max2(hi, val);
            acc[0] += val;
            acc[1] += val * val;
        }
    }
    // block-level min / max (order-independent), then one slot per block; the sums go through grid_sum
    __shared__ double s_lo[32], s_hi[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min2(lo, __shfl_down_sync(0xffffffffu, lo, o));
        hi = max2(hi, __shfl_down_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) s_lo[warp] = lo, s_hi[warp] = hi;
    __syncthreads();
    const unsigned int nblk = gridDim.x * gridDim.y, bid = blockIdx.x + blockIdx.y * gridDim.x;
    double *pmin = partials + 2 * (size_t)nblk, *pmax = pmin + nblk;
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) lo = min2(lo, s_lo[w]), hi = max2(hi, s_hi[w]);
        pmin[bid] = lo, pmax[bid] = hi;
    }
    grid_sum<2>(acc, partials, &sc->ticket[1], [=](double(&t)[2]) { // the last block: every partial is visible here
        double a = 1e30, b = -1e30;
        for (unsigned int k = 0; k < nblk; ++k) a = min2(a, __ldcg(&pmin[k])), b = max2(b, __ldcg(&pmax[k]));
        sc->viz[0] = a, sc->viz[1] = b, sc->viz[2] = t[0], sc->viz[3] = t[1];
    });
}

// ---- Colormaps::apply_colormap (colormaps.cpp) ---------------------------------------------------------------
__constant__ float VIZ_HORNER[2][7][3] = { // degree-6 fits, lowest order first: viridis, plasma (colormaps.cpp:8-14, :27-33)
    {{0.2777273272234177f, 0.005407344544966578f, 0.3340998053353061f},
     {0.1050930431085774f, 1.404613529898575f, 1.384590162594685f},
     {-0.3308618287255563f, 0.214847559468213f, 0.09509516302823659f},
     {-4.634230498983486f, -5.799100973351585f, -19.33244095627987f},
     {6.228269936347081f, 14.17993336680509f, 56.69055260068105f},
     {4.776384997670288f, -13.74514537774601f, -65.35303263337234f},
     {-5.435455855934631f, 12.86216349749025f, 26.312873249486f}},
    {{0.05873234392399702f, 0.02333670892565664f, 0.5433401826748754f},
     {2.176514634195958f, 0.2383834171260182f, 0.7539604599784036f},
     {-2.689460476458034f, -7.455851135738909f, 3.110799939717086f},
     {6.130348345893603f, 42.3461881477227f, -28.51885465332158f},
     {-11.10743619062271f, -82.66631104428044f, 60.13984767418263f},
     {10.02306557647065f, 71.41361770095349f, -54.07218655560067f},
     {-3.658713842777788f, -22.93153465461149f, 18.19190778539828f}}};
__constant__ float VIZ_TURBO[5][3] = {{0.18995f, 0.07176f, 0.23217f}, {0.30315f, 0.50427f, 0.94791f},
                                      {0.38343f, 0.87344f, 0.54467f}, {0.98853f, 0.64411f, 0.09395f},
                                      {0.95839f, 0.82344f, 0.08425f}};

__device__ __forceinline__ float viz_clamp01(float x) { return (x < 0.0f) ? 0.0f : ((1.0f < x) ? 1.0f : x); }

__device__ __forceinline__ void viz_colormap(int cm, float t, float (&c)[3]) {
    if (cm == 2) { // jet
        if (t < 0.125f) { c[0] = 0.0f; c[1] = 0.0f; c[2] = 0.5f + 4.0f * t; }
        else if (t < 0.375f) { c[0] = 0.0f; c[1] = 4.0f * (t - 0.125f); c[2] = 1.0f; }
        else if (t < 0.625f) { c[0] = 4.0f * (t - 0.375f); c[1] = 1.0f; c[2] = 1.0f - 4.0f * (t - 0.375f); }
        else if (t < 0.875f) { c[0] = 1.0f; c[1] = 1.0f - 4.0f * (t - 0.625f); c[2] = 0.0f; }
        else { c[0] = 1.0f - 4.0f * (t - 0.875f); c[1] = 0.0f; c[2] = 0.0f; }
    } else if (cm == 3) { // grayscale
        c[0] = c[1] = c[2] = t;
    } else if (cm == 4) { // turbo: quartic in power form
        const float t2 = t * t, t3 = t2 * t, t4 = t3 * t;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            c[k] = viz_clamp01(VIZ_TURBO[0][k] + VIZ_TURBO[1][k] * t + VIZ_TURBO[2][k] * t2 + VIZ_TURBO[3][k] * t3 + VIZ_TURBO[4][k] * t4);
    } else if (cm == 5) { // red-yellow-blue
        if (t < 0.5f) { c[0] = 1.0f; c[1] = 2.0f * t; c[2] = 0.0f; }
        else { const float t2 = 2.0f * (t - 0.5f); c[0] = 1.0f - t2; c[1] = 1.0f - t2; c[2] = t2; }
    } else { // viridis (0 and every unknown id, colormaps.cpp:124), plasma (1)
        const int m = cm == 1 ? 1 : 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float acc = VIZ_HORNER[m][6][k];
#pragma unroll
            for (int d = 5; d >= 1; --d) acc = VIZ_HORNER[m][d][k] + t * acc;
            c[k] = viz_clamp01(VIZ_HORNER[m][0][k] + t * acc);
        }
    }
}

// gui.cpp:1534-1551: normalise, gamma, colour, 3 bytes per cell
__global__ void __launch_bounds__(256)
    k_viz_rgb(size_t n, const double *__restrict__ field, double vmin, double scale, int normalize, double inv_gamma,
              int colormap, unsigned char *__restrict__ rgb) { pdl_begin();
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        double val = field[k];
        if (normalize) val = fabs(val);
        double norm = clamp3((val - vmin) * scale, 0.0, 1.0);
        norm = pow(norm, inv_gamma);
        float c[3];
        viz_colormap(colormap, static_cast<float>(norm), c);
        rgb[3 * k + 0] = static_cast<unsigned char>(c[0] * 255.0f);
        rgb[3 * k + 1] = static_cast<unsigned char>(c[1] * 255.0f);
        rgb[3 * k + 2] = static_cast<unsigned char>(c[2] * 255.0f);
    }
}

// shapes.cpp:33-71 for ONE shape: every cell of the (host-clamped) box that lies inside the shape gets its left u face
// and its bottom v face zeroed (solid) or scaled by 0.1 (porous).  One thread per cell; a cell owns those two faces,
// so there are no write conflicts inside a shape; shapes are applied in order, one launch each.
// (row slabs: j0 .. j1 are the rows of the box that fall into this slab, in GLOBAL cell rows; g.j0 maps them to local rows)
__global__ void __launch_bounds__(256)
    k_shape_apply(Geom g, double *u, double *v, int i0, int j0, int i1, int j1, int circle, float cx, float cy, float rad,
                  int porous) { pdl_begin();
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x, j = j0 + blockIdx.y;
    if (i > i1 || j > j1) return;
    if (circle) { // cell centre within the radius, all in float (shapes.cpp:41-47)
        const float dx = (static_cast<float>(i) + 0.5f) - cx, dy = (static_cast<float>(j) + 0.5f) - cy;
        if (!(dx * dx + dy * dy <= rad * rad)) return;
    }
    const size_t o = gidx(g, i + 1, j - g.j0 + 1);
    if (porous) {
        u[o] *= 0.1;
        v[o] *= 0.1;
    } else {
        u[o] = 0.0;
        v[o] = 0.0;
    }
}
