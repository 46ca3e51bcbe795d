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


# ------------------------------------------------------------------------------------------------------------------
# Round 2: the one-way sends of a time step (k_halo_send + waits in the consumers' edge blocks, capi.cu enqueue_step).
# No handshake: a send into halo buffer X is legal because a reduction (a barrier) lies between the neighbour's last
# read of X and the send.  The model checks exactly that: every writer asserts that the previous content of the slot it
# overwrites has been consumed, every consumer that it reads this step's row of the right buffer -- with random delays,
# ranks running ahead as far as the two barriers of a step let them, and the data flag shared by all buffers of a side.
# ------------------------------------------------------------------------------------------------------------------
BUFFERS = ["u1v1", "vstar", "rhs", "p_alt", "p_row", "uvp"]  # order of the sends inside one mg step


class SlabBlock(Block):
    def __init__(self, n):
        super().__init__(n)
        self.rows = {b: [None, None] for b in BUFFERS}       # halo rows per buffer and side
        self.consumed = {b: [0, 0] for b in BUFFERS}          # step whose content the owner has finished reading


def barrier(r, n, blocks, me, red_epoch, rng):
    par = red_epoch & 1
    for t in range(n):
        jitter(rng)
        blocks[t].red_flag[par][r] = red_epoch
    for t in range(n):
        spin(lambda: me.red_flag[par][t] >= red_epoch, f"rank {r} barrier {red_epoch}")


def slab_rank(r, n, blocks, steps, errs, seed):
    rng = random.Random(seed * 977 + r)
    me = blocks[r]
    lower, upper = (r - 1 if r > 0 else None), (r + 1 if r < n - 1 else None)
    halo_epoch = red_epoch = 0
    pending = {}  # buffer -> epoch its consumer waits for

    def send(buf, step):
        nonlocal halo_epoch
        halo_epoch += 1
        for nb, slot in ((lower, 1), (upper, 0)):
            if nb is not None:
                jitter(rng)
                # the handshake-free send is only legal if the neighbour is done with what the slot holds
                assert blocks[nb].consumed[buf][slot] == step - 1, (r, step, buf, blocks[nb].consumed[buf][slot])
                blocks[nb].rows[buf][slot] = (r, step, buf)
                blocks[nb].data[slot] = halo_epoch       # one flag per side for all buffers: epochs only grow
        pending[buf] = halo_epoch

    def consume(buf, step):
        ep = pending.pop(buf)
        for nb, mine in ((lower, 0), (upper, 1)):
            if nb is not None:
                spin(lambda: me.data[mine] >= ep, f"rank {r} step {step} waits for {buf}")
                jitter(rng)
                assert me.rows[buf][mine] == (nb, step, buf), (r, step, buf, me.rows[buf][mine])
                me.consumed[buf][mine] = step

    try:
        for step in range(1, steps + 1):
            if step > 1:
                consume("uvp", step - 1)                 # stage 1 reads what the neighbours sent at the end of the last step
            send("u1v1", step); consume("u1v1", step)    # stage 2's edge strips
            send("vstar", step)
            send("rhs", step)
            consume("vstar", step)                       # divergence, top block (scheduled last)
            consume("rhs", step)                         # smoother 1, edge strips
            red_epoch += 1; barrier(r, n, blocks, me, red_epoch, rng)   # residual of smooth() 1
            send("p_alt", step); consume("p_alt", step)  # smoother 2
            send("p_row", step); consume("p_row", step)  # projection (the last residual rides on the next barrier)
            red_epoch += 1; barrier(r, n, blocks, me, red_epoch, rng)   # end of the step
            send("uvp", step)
    except BaseException as ex:  # noqa: BLE001
        errs.append((r, repr(ex)))


def run_steps(n, steps, seed):
    blocks = [SlabBlock(n) for _ in range(n)]
    errs = []
    ths = [threading.Thread(target=slab_rank, args=(r, n, blocks, steps, errs, seed)) for r in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
    assert not errs, errs


def test_one_way_sends_of_a_step_two_ranks():
    run_steps(2, 120, seed=5)


def test_one_way_sends_of_a_step_four_and_eight_ranks():
    run_steps(4, 60, seed=6)
    run_steps(8, 30, seed=7)


def test_missing_barrier_is_caught_by_the_model():
    """The check has teeth: sending the smoother's rows twice without a barrier in between (what capi.cu's
    sent_since_barrier bookkeeping turns into a handshake exchange) violates the invariant."""
    n, blocks, errs = 2, [SlabBlock(2) for _ in range(2)], []

    def bad(r):
        try:
            me, nb, slot = blocks[r], 1 - r, (1 if r == 1 else 0)
            for step in (1, 2):  # second send while the neighbour has not consumed the first
                assert blocks[nb].consumed["p_alt"][slot] == step - 1
                blocks[nb].rows["p_alt"][slot] = (r, step, "p_alt")
        except AssertionError as ex:
            errs.append(r)

    ths = [threading.Thread(target=bad, args=(r,)) for r in range(n)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert sorted(errs) == [0, 1]
