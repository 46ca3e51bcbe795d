/*
 * cfd_vcycle.c -- CPU restatement of the TRUE geometric V-cycle the B200 path offers as an extension
 * (mg_mode = CFD_B200_MG_VCYCLE).  TEST INFRASTRUCTURE ONLY, like everything under oracle/.
 *
 * PARITY UNPINNED: the reference has no multigrid code at all -- its "Multigrid" is fine-grid red-black Gauss-Seidel
 * (pressure.cpp:113-129) and `mg_levels` / `rbgs_sweeps` (pressure.hpp:10-12, parsed at main.cpp:75-78) are never
 * read.  There is therefore no reference output to pin this file on.  It defines the algorithm the CUDA kernels must
 * reproduce bit for bit, and is itself validated by what multigrid must do: reduce the residual of L p = rhs by a
 * grid-independent factor per cycle (tests/test_vcycle.py).
 *
 * Algorithm (cell-centred 5-point Laplacian, p = 0 on the wall faces -- NOT the PCG path's zero-ghost operator, whose
 * boundary would sit at a different place on every level):
 *   level l has nx/2^l x ny/2^l cells; levels actually used = min(mg_levels, as long as both sizes stay even and >= 4)
 *   V(l):  rbgs_sweeps red-black Gauss-Seidel sweeps  p = ((pr+pl)*idx2 + (pt+pb)*idy2 - rhs) / (2*(idx2+idy2) + wall terms)
 *            (the sign the reference's smoother gets wrong, SURVEY.md fact 2, is fixed here);
 *          coarsest level: 16 x rbgs_sweeps sweeps instead; otherwise
 *          r = rhs - L p ; rhs_{l+1} = mean of the 4 children of each coarse cell ; p_{l+1} = 0 ; V(l+1) ;
 *          p += bilinear interpolation of p_{l+1} (9/16, 3/16, 3/16, 1/16) ; rbgs_sweeps sweeps again.
 *   after every cycle res = ||rhs - L p||_2 ; stop when res < tol or after `vcycles` cycles.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int nx, ny, P;
    double idx2, idy2, denom;
    double *p, *rhs;
} level_t;

/* Boundary: p = 0 on the wall FACE (mirror ghost = -interior), the same physical place on every level.  The ghost
 * cells in memory stay 0; a cell touching a wall instead gets idx2 (idy2) added to its diagonal per wall side:
 * bx = number of x-walls the cell touches (0, 1, or 2 when nx == 1), by likewise. */
static inline double wall_diag(const level_t *L, int ii, int jj) {
    int bx = (ii == 1) + (ii == L->nx), by = (jj == 1) + (jj == L->ny);
    return (double)bx * L->idx2 + (double)by * L->idy2;
}
static void smooth(level_t *L, int sweeps) {
    const int P = L->P;
    for (int s = 0; s < sweeps; ++s)
        for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for schedule(static)
            for (int jj = 1; jj <= L->ny; ++jj)
                for (int ii = 1 + ((jj - 1 + colour) & 1); ii <= L->nx; ii += 2) {
                    double *q = L->p + (size_t)jj * P + ii;
                    double sum = (q[1] + q[-1]) * L->idx2 + (q[P] + q[-P]) * L->idy2;
                    q[0] = (sum - L->rhs[(size_t)jj * P + ii]) / (L->denom + wall_diag(L, ii, jj));
                }
        }
}
static inline double resid(const level_t *L, int ii, int jj) {
    const int P = L->P;
    const double *q = L->p + (size_t)jj * P + ii;
    double Ap = (q[1] - 2.0 * q[0] + q[-1]) * L->idx2 + (q[P] - 2.0 * q[0] + q[-P]) * L->idy2;
    Ap = Ap - wall_diag(L, ii, jj) * q[0];
    return L->rhs[(size_t)jj * P + ii] - Ap;
}
static void restrict_residual(const level_t *F, level_t *C) {
#pragma omp parallel for schedule(static)
    for (int J = 1; J <= C->ny; ++J)
        for (int I = 1; I <= C->nx; ++I) {
            int ii = 2 * I - 1, jj = 2 * J - 1;
            double a = resid(F, ii, jj), b = resid(F, ii + 1, jj), c = resid(F, ii, jj + 1), d = resid(F, ii + 1, jj + 1);
            C->rhs[(size_t)J * C->P + I] = 0.25 * ((a + b) + (c + d));
            C->p[(size_t)J * C->P + I] = 0.0;
        }
}
/* Bilinear (cell-centred) prolongation: parent 9/16, the two nearer coarse neighbours 3/16 each, the diagonal one
 * 1/16; coarse ghosts are 0.  (Piecewise-constant prolongation with averaging restriction violates the order rule
 * m_P + m_R > 2 and diverges on deep hierarchies -- measured, see tests/test_vcycle.py.) */
static void prolong_add(level_t *F, const level_t *C) {
#pragma omp parallel for schedule(static)
    for (int jj = 1; jj <= F->ny; ++jj)
        for (int ii = 1; ii <= F->nx; ++ii) {
            int I = (ii + 1) / 2, J = (jj + 1) / 2;
            int di = (ii & 1) ? -1 : 1, dj = (jj & 1) ? -1 : 1; /* odd raw index = left/lower child */
            const double *c = C->p + (size_t)J * C->P + I;
            double e = 0.0625 * ((9.0 * c[0] + 3.0 * (c[di] + c[dj * C->P])) + c[dj * C->P + di]);
            F->p[(size_t)jj * F->P + ii] += e;
        }
}
static void vcycle(level_t *lv, int l, int nlev, int sweeps) {
    if (l == nlev - 1) {
        smooth(&lv[l], nlev == 1 ? sweeps : 16 * sweeps);
        return;
    }
    smooth(&lv[l], sweeps);
    restrict_residual(&lv[l], &lv[l + 1]);
    vcycle(lv, l + 1, nlev, sweeps);
    prolong_add(&lv[l], &lv[l + 1]);
    smooth(&lv[l], sweeps);
}

int cfdo_vcycle_levels(int nx, int ny, int mg_levels) {
    int n = 1;
    while (n < mg_levels && nx % 2 == 0 && ny % 2 == 0 && nx / 2 >= 4 && ny / 2 >= 4) nx /= 2, ny /= 2, ++n;
    return n;
}

/* p, rhs: reference host layout (pitch nx+2, rows ny+2); the ghost ring of p is forced to 0 (Dirichlet data). */
double cfdo_vcycle_solve(int nx, int ny, double Lx, double Ly, double *p, const double *rhs, int mg_levels,
                         int vcycles, int rbgs_sweeps, double tol, int *cycles_done) {
    int nlev = cfdo_vcycle_levels(nx, ny, mg_levels);
    level_t *lv = calloc((size_t)nlev, sizeof(level_t));
    for (int l = 0; l < nlev; ++l) {
        lv[l].nx = nx >> l;
        lv[l].ny = ny >> l;
        lv[l].P = lv[l].nx + 2;
        double dx = Lx / (double)lv[l].nx, dy = Ly / (double)lv[l].ny;
        lv[l].idx2 = 1.0 / (dx * dx);
        lv[l].idy2 = 1.0 / (dy * dy);
        lv[l].denom = 2.0 * (lv[l].idx2 + lv[l].idy2);
        size_t n = (size_t)lv[l].P * (lv[l].ny + 2);
        if (l == 0) {
            lv[l].p = p;
            lv[l].rhs = (double *)rhs;
        } else {
            lv[l].p = calloc(n, sizeof(double));
            lv[l].rhs = calloc(n, sizeof(double));
        }
    }
    /* zero ghosts on the fine level */
    for (int ii = 0; ii <= nx + 1; ++ii) p[ii] = 0.0, p[(size_t)(ny + 1) * lv[0].P + ii] = 0.0;
    for (int jj = 0; jj <= ny + 1; ++jj) p[(size_t)jj * lv[0].P] = 0.0, p[(size_t)jj * lv[0].P + nx + 1] = 0.0;
    double res = 1e9;
    int done = 0;
    for (int c = 0; c < vcycles; ++c) {
        vcycle(lv, 0, nlev, rbgs_sweeps);
        double sum = 0.0;
        for (int jj = 1; jj <= ny; ++jj)
            for (int ii = 1; ii <= nx; ++ii) {
                double r = resid(&lv[0], ii, jj);
                sum += r * r;
            }
        res = sqrt(sum);
        done = c + 1;
        if (res < tol) break;
    }
    for (int l = 1; l < nlev; ++l) free(lv[l].p), free(lv[l].rhs);
    free(lv);
    if (cycles_done) *cycles_done = done;
    return isfinite(res) ? res : 1e9;
}
