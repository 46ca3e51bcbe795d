"""CPU thread model of the exchange protocol of k_pcg_cluster (macos-cfd_b200/csrc/kernels_pcg_cluster.cuh).

Every CTA of the cluster stores its partial sums into slot [rank] of ALL CTAs (slotA before the barrier that ends phase A,
slotB before the one that ends phase B) and its boundary r rows into its neighbours' halo buffers (phase B), without any
handshake: the two cluster barriers of an iteration are the only synchronisation, and no buffer is double-buffered.  The model
runs the same sequence with threads and a reusable barrier, random delays in every phase, and checks that
  * every value a CTA reads after a barrier is the one written for exactly this iteration and phase (never a stale one,
    never one of a later iteration: a fast CTA must not be able to overwrite what a slow one has not read yet);
  * all CTAs take the same decisions (they add the same 16 partials in the same order).
It models the ORDER of the accesses, not the memory model: visibility is what barrier.cluster's release / acquire gives.
"""
import random
import threading

import pytest


def run_model(nctas, iters, seed, stop_after=None):
    rng = random.Random(seed)
    delays = [[rng.random() * 2e-4 for _ in range(4 * (iters + 1))] for _ in range(nctas)]
    bar = threading.Barrier(nctas)
    slotA = [[None] * nctas for _ in range(nctas)]   # slotA[owner][writer]
    slotB = [[None] * nctas for _ in range(nctas)]
    rhalo = [[None, None] for _ in range(nctas)]     # [0] written by the CTA below, [1] by the one above
    errors, decisions = [], [[] for _ in range(nctas)]

    def nap(b, k):
        d = delays[b][k]
        if d > 1e-4:
            threading.Event().wait(d)

    def send_rows(b, tag):
        if b > 0:
            rhalo[b - 1][1] = (tag, b)
        if b + 1 < nctas:
            rhalo[b + 1][0] = (tag, b)

    def cta(b):
        try:
            k = 0
            # start: r rows + partials (slotB), barrier, totals
            send_rows(b, ("init",))
            for o in range(nctas):
                slotB[o][b] = (("init",), b)
            nap(b, k); k += 1
            bar.wait()
            assert all(slotB[b][w] == (("init",), w) for w in range(nctas)), ("init totals", b, slotB[b])
            last = ("init",)
            for it in range(iters):
                # phase A: read the neighbours' r rows of the previous phase B (or of the start)
                nap(b, k); k += 1
                if b > 0:
                    assert rhalo[b][0] == (last, b - 1), ("halo below", b, it, rhalo[b][0])
                if b + 1 < nctas:
                    assert rhalo[b][1] == (last, b + 1), ("halo above", b, it, rhalo[b][1])
                for o in range(nctas):
                    slotA[o][b] = (("A", it), b)
                nap(b, k); k += 1
                bar.wait()
                got = [slotA[b][w] for w in range(nctas)]
                assert got == [(("A", it), w) for w in range(nctas)], ("slotA", b, it, got)
                decisions[b].append(("A", it, tuple(g[0] for g in got)))
                if stop_after == ("A", it):  # e.g. a breakdown or a non-finite flag: everybody leaves at the same barrier
                    break
                # phase B: own update, boundary rows to the neighbours, partials
                nap(b, k); k += 1
                send_rows(b, ("B", it))
                for o in range(nctas):
                    slotB[o][b] = (("B", it), b)
                nap(b, k); k += 1
                bar.wait()
                got = [slotB[b][w] for w in range(nctas)]
                assert got == [(("B", it), w) for w in range(nctas)], ("slotB", b, it, got)
                decisions[b].append(("B", it, tuple(g[0] for g in got)))
                last = ("B", it)
                if stop_after == ("B", it):
                    break
            bar.wait()  # nobody leaves while a peer may still store into its shared memory
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
            bar.abort()

    ts = [threading.Thread(target=cta, args=(b,)) for b in range(nctas)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    return errors, decisions


@pytest.mark.parametrize("nctas,iters,seed,stop", [(2, 6, 1, None), (4, 8, 2, None), (16, 6, 3, None), (16, 5, 4, ("A", 3)),
                                                   (8, 5, 5, ("B", 2))])
def test_cluster_exchange_needs_no_double_buffering(nctas, iters, seed, stop):
    errors, decisions = run_model(nctas, iters, seed, stop)
    assert not errors, errors[:3]
    assert all(d == decisions[0] for d in decisions), "the CTAs took different decisions"
    assert decisions[0], "no phase ran"
