"""A thread model of the flag protocol of macos-cfd_b200/csrc/p2p.cuh (the peer-memory transport of the row slabs).

The CUDA kernels cannot run here (no GPU), but what makes them correct is a protocol, not arithmetic, and that can be
exercised on the CPU: N "ranks" (threads) with one "block" each (flags + slots, written by the peers, read by the
owner), random delays everywhere, many epochs.  Checked:
  * k_p2p_reduce: epoch-stamped slots double-buffered on the epoch parity -- no rank ever reads a slot of another epoch,
    all ranks compute identical totals, although nothing is ever reset and a fast rank may be one reduction ahead;
  * k_halo_xchg: enter flag -> wait -> store rows -> data flag -> wait: a rank's halo row is never overwritten while it
    may still read the previous one, and what it reads after the wait is the neighbour's row of THIS exchange;
  * the merged form (rows riding on a reduction, no enter flag) is safe because the reduction is a barrier.
The real thing is covered by tests/test_gpu_multi.py on 2 and 4 GPUs (bit-identical to one GPU, both transports)."""
import random
import threading
import time


class Block:
    def __init__(self, n):
        self.enter = [0, 0]          # [0] written by the rank below, [1] by the rank above
        self.data = [0, 0]
        self.red_flag = [[0] * n, [0] * n]
        self.red_val = [[None] * n, [None] * n]
        self.halo = [None, None]     # "halo rows": [0] from below, [1] from above


def jitter(rng):
    if rng.random() < 0.3:
        time.sleep(rng.random() * 2e-4)


def spin(cond, what):
    t0 = time.time()
    while not cond():
        if time.time() - t0 > 20:
            raise TimeoutError(what)
        time.sleep(0)


def rank_main(r, n, blocks, epochs, merged_every, out, seed):
    rng = random.Random(seed * 1000 + r)
    me = blocks[r]
    red_epoch = halo_epoch = 0
    lower, upper = (r - 1 if r > 0 else None), (r + 1 if r < n - 1 else None)
    log = []
    for e in range(1, epochs + 1):
        row = (r, e)                                   # what this rank's edge row holds in this round
        # ---- a plain exchange (k_halo_xchg): handshake, rows, data flag
        halo_epoch += 1
        for nb, slot in ((lower, 1), (upper, 0)):
            if nb is not None:
                jitter(rng)
                blocks[nb].enter[slot] = halo_epoch
        for nb, mine in ((lower, 0), (upper, 1)):
            if nb is not None:
                spin(lambda: me.enter[mine] >= halo_epoch, f"rank {r} enter {halo_epoch}")
        for nb, slot in ((lower, 1), (upper, 0)):
            if nb is not None:
                jitter(rng)
                blocks[nb].halo[slot] = row            # the neighbour said it is past every reader of that row
                blocks[nb].data[slot] = halo_epoch
        for nb, mine in ((lower, 0), (upper, 1)):
            if nb is not None:
                spin(lambda: me.data[mine] >= halo_epoch, f"rank {r} data {halo_epoch}")
                jitter(rng)                            # "consumer kernels" read the halo row now
                assert me.halo[mine] == (nb, e), (r, e, me.halo[mine])
        # ---- a reduction (k_p2p_reduce), every merged_every-th one with rows riding on it (no enter flag)
        red_epoch += 1
        par = red_epoch & 1
        val = (r + 1) * e
        for t in range(n):
            jitter(rng)
            blocks[t].red_val[par][r] = (red_epoch, val)
            blocks[t].red_flag[par][r] = red_epoch
        for t in range(n):
            spin(lambda: me.red_flag[par][t] >= red_epoch, f"rank {r} red {red_epoch}")
        jitter(rng)
        got = [me.red_val[par][t] for t in range(n)]
        assert all(g[0] == red_epoch for g in got), (r, red_epoch, got)   # never a slot of another epoch
        log.append(sum(g[1] for g in got))
        if e % merged_every == 0:
            halo_epoch += 1
            row2 = (r, -e)
            for nb, slot in ((lower, 1), (upper, 0)):
                if nb is not None:
                    jitter(rng)
                    blocks[nb].halo[slot] = row2       # safe: everybody has contributed, so nobody still reads `row`
                    blocks[nb].data[slot] = halo_epoch
            for nb, mine in ((lower, 0), (upper, 1)):
                if nb is not None:
                    spin(lambda: me.data[mine] >= halo_epoch, f"rank {r} merged data {halo_epoch}")
                    jitter(rng)
                    assert me.halo[mine] == (nb, -e), (r, e, me.halo[mine])
    out[r] = log


def run(n, epochs, merged_every, seed):
    blocks = [Block(n) for _ in range(n)]
    out, errs = [None] * n, []

    def guard(r):
        try:
            rank_main(r, n, blocks, epochs, merged_every, out, seed)
        except BaseException as ex:  # noqa: BLE001 -- reported below
            errs.append((r, repr(ex)))

    ths = [threading.Thread(target=guard, args=(r,)) for r in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
    assert not errs, errs
    assert all(o == out[0] for o in out)               # identical totals on every rank, every epoch
    assert out[0] == [sum((r + 1) * e for r in range(n)) for e in range(1, epochs + 1)]


def test_protocol_two_ranks():
    run(2, 150, 3, seed=1)


def test_protocol_middle_ranks_have_two_neighbours():
    run(4, 120, 2, seed=2)


def test_protocol_eight_ranks():
    run(8, 60, 1, seed=3)
