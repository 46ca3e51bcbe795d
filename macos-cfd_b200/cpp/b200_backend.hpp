// b200_backend.hpp -- what sits between the reference-shaped classes and the C ABI (include/cfd_b200.h).
//
// One Session per Grid object holds the device State.  The caller owns the HOST State (plain Field2D arrays, as in
// the reference) and may read or write it between steps (the GUI does); two policies keep that legal:
//
//   Coherence::Eager (default; always correct, a true drop-in): every TimeIntegrator::step uploads s.u, s.v, s.p
//       first and downloads s.u, s.v, s.p, s.rhs afterwards -- the host State is exactly what the reference leaves.
//   Coherence::Lazy  (headless runs): the State lives on the device.  step() uploads only after
//       b200::mark_host_dirty(), and the host arrays are refreshed only by b200::sync_to_host().  The diagnostics
//       divergence_l2 / max_velocity called on the State's own u, v are answered from the device without a copy.
//
// Select with b200::set_coherence() or the environment variable CFD_B200_COHERENCE=eager|lazy.
#pragma once
#include "solver/state.hpp"
#include "solver/grid.hpp"

struct cfd_b200_solver;

namespace b200 {

enum class Coherence { Eager, Lazy };
void set_coherence(Coherence c);
Coherence coherence();

// Lazy mode: the caller changed the host State / wants the host State to be current.
void mark_host_dirty(const State &s);
void sync_to_host(State &s);

// CUDA device the next Session is created on (default 0, or CFD_B200_DEVICE).
void set_device(int device);

// What the last step logged in the reference (stderr) and more, for drivers that want it.
struct StepReport {
    double dt = 0, residual = 0, div_before = 0, div_after = 0;
    int iters = 0, pcg_breakdown = -1, nonfinite = 0;
};
StepReport last_report(const Grid &g);

struct Session; // opaque
Session *acquire(const Grid &g);
void release(Session *s);
cfd_b200_solver *handle(Session *s);

} // namespace b200
