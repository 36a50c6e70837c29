"""Exhaustive interleaving check of the two peer-memory exchange protocols of a sharded sweep (ig_kernels.cu
peer_allreduce_kernel, p_peer_signal_kernel + p_dirichlet_kernel<PEER>).  No GPU: the protocols are
restated as per-rank step sequences over shared cells and every interleaving of the ranks' steps is explored (memoised
depth-first search over (program counters, memory)); the assertions are the hazards the design has to exclude:

  scalar all-reduce   slots alternate by the parity of the sequence number.  A rank must never read a slot that holds another
                      round's value, although a fast rank may already be storing the NEXT round while a slow one still reads.
  tally -> P          one tally buffer and two P buffers per rank.  A rank pulls every rank's tally only when it is final,
                      pushes its block into every rank's NEXT P buffer only when nobody reads that buffer any more, clears
                      its own tally only when every rank has pulled it, and the next sweep kernel sees a complete P.

A step is atomic here; on the device the order inside a rank is given by the system-scope fences and the stream order."""
import sys


def _explore(nranks, programs, init_mem, step):
    """programs[r] = list of ops; step(mem, r, op) -> new mem, or None when the op has to wait.  Returns #states."""
    sys.setrecursionlimit(100000)
    seen = set()
    stack = [((0,) * nranks, init_mem)]
    finished = False
    while stack:
        pcs, mem = stack.pop()
        if (pcs, mem) in seen:
            continue
        seen.add((pcs, mem))
        moved = False
        done = True
        for r in range(nranks):
            if pcs[r] >= len(programs[r]):
                continue
            done = False
            new = step(mem, r, programs[r][pcs[r]])
            if new is None:
                continue
            moved = True
            stack.append((pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:], new))
        if done:
            finished = True
        else:
            assert moved, f"deadlock at {pcs}"
    assert finished
    return len(seen)


# ---------------------------------------------------------------------------------------------------------------------
# scalar all-reduce: memory = (slots, flags); slots[owner][parity][src] = round stored by src, flags likewise
# ---------------------------------------------------------------------------------------------------------------------
def _allreduce_program(W, rounds):
    prog = []
    for seq in range(1, rounds + 1):
        prog += [("store", p, seq) for p in range(W)]          # my sums into my slot of every rank's buffer
        prog += [("flag", p, seq) for p in range(W)]           # (fence) then my sequence number behind them
        prog += [("wait", seq)]                                # all W numbers have arrived in MY buffer
        prog += [("read", q, seq) for q in range(W)]           # add the W slots
    return prog


def _allreduce_step(W):
    def idx(owner, par, src):
        return (owner * 2 + par) * W + src

    def step(mem, r, op):
        slots, flags = mem
        if op[0] == "store":
            _, p, seq = op
            s = list(slots); s[idx(p, seq & 1, r)] = seq
            return (tuple(s), flags)
        if op[0] == "flag":
            _, p, seq = op
            f = list(flags); f[idx(p, seq & 1, r)] = seq
            return (slots, tuple(f))
        if op[0] == "wait":
            seq = op[1]
            return mem if all(flags[idx(r, seq & 1, q)] == seq for q in range(W)) else None
        _, q, seq = op
        assert slots[idx(r, seq & 1, q)] == seq, f"rank {r} reads round {slots[idx(r, seq & 1, q)]} of rank {q} in round {seq}"
        return mem
    return step


def test_scalar_allreduce_slots_are_never_read_stale_or_overwritten_early():
    for W, rounds in ((2, 4), (3, 3)):
        zero = (0,) * (W * 2 * W)
        n = _explore(W, [_allreduce_program(W, rounds)] * W, (zero, zero), _allreduce_step(W))
        assert n > 100


def test_a_single_slot_per_rank_would_be_unsafe():
    """The same protocol WITHOUT the parity alternation has an interleaving in which a fast rank's next round lands in a slot a
    slow rank has not read yet: the check has teeth."""
    W, rounds = 2, 2

    def step(mem, r, op):
        slots, flags = mem
        if op[0] == "store":
            s = list(slots); s[op[1] * W + r] = op[2]
            return (tuple(s), flags)
        if op[0] == "flag":
            f = list(flags); f[op[1] * W + r] = op[2]
            return (slots, tuple(f))
        if op[0] == "wait":
            return mem if all(flags[r * W + q] >= op[1] for q in range(W)) else None
        if slots[r * W + op[1]] != op[2]:
            raise RuntimeError("stale")
        return mem
    zero = (0,) * (W * W)
    try:
        _explore(W, [_allreduce_program(W, rounds)] * W, (zero, zero), step)
    except RuntimeError:
        return
    raise AssertionError("expected a hazard without double buffering")


# ---------------------------------------------------------------------------------------------------------------------
# tally -> P.  Per rank and sweep t (the two streams of a rank are ordered by events, so a rank's steps are sequential):
#   zq(t)     reads P buffer t & 1 (every block must be version t), writes its tally := t (the buffer must be clear)
#   signal 0  "my tally is final": flag0[p][me] = t for every p; wait for flag0[me][q] >= t
#   pull      read every rank's tally: must be version t
#   push      write block `me` of every rank's P buffer (t + 1) & 1 := t + 1; that rank must not be reading it (its zq(t - 1),
#             which used this buffer, is over: it signalled tally(t))
#   signal 1  "my block is in your buffer": flag1[p][me] = t; wait for flag1[me][q] >= t
#   clear     own tally := 0; every rank must have pulled it
# memory = (tally[W], pulled[W][W] ghost, P[W][2][W], flag0[W][W], flag1[W][W], reading[W] ghost: buffer being read or -1)
# ---------------------------------------------------------------------------------------------------------------------
def _p_program(W, sweeps):
    prog = []
    for t in range(1, sweeps + 1):
        prog += [("zq_begin", t), ("zq_end", t)]
        prog += [("f0", p, t) for p in range(W)] + [("w0", t)]
        prog += [("pull", q, t) for q in range(W)]
        prog += [("push", p, t) for p in range(W)]
        prog += [("f1", p, t) for p in range(W)] + [("w1", t)]
        prog += [("clear", t)]
    return prog


def _p_step(W):
    def step(mem, r, op):
        tally, pulled, P, f0, f1, reading = mem
        k = op[0]
        if k == "zq_begin":
            t = op[1]
            buf = t & 1
            assert all(P[(r * 2 + buf) * W + b] == t for b in range(W)), f"rank {r} sweep {t}: P incomplete"
            assert tally[r] == 0, f"rank {r} sweep {t}: tally not cleared"
            rd = list(reading); rd[r] = buf
            return (tally, pulled, P, f0, f1, tuple(rd))
        if k == "zq_end":
            t = op[1]
            ta = list(tally); ta[r] = t
            rd = list(reading); rd[r] = -1
            return (tuple(ta), pulled, P, f0, f1, tuple(rd))
        if k in ("f0", "f1"):
            _, p, t = op
            f = list(f0 if k == "f0" else f1); f[p * W + r] = t
            return (tally, pulled, P, tuple(f), f1, reading) if k == "f0" else (tally, pulled, P, f0, tuple(f), reading)
        if k in ("w0", "w1"):
            f = f0 if k == "w0" else f1
            return mem if all(f[r * W + q] >= op[1] for q in range(W)) else None
        if k == "pull":
            _, q, t = op
            assert tally[q] == t, f"rank {r} pulls tally {tally[q]} of rank {q} in sweep {t}"
            pl = list(pulled); pl[q * W + r] = t
            return (tally, tuple(pl), P, f0, f1, reading)
        if k == "push":
            _, p, t = op
            buf = (t + 1) & 1
            assert reading[p] != buf, f"rank {r} overwrites the P buffer rank {p} is reading (sweep {t})"
            Pn = list(P); Pn[(p * 2 + buf) * W + r] = t + 1
            return (tally, pulled, tuple(Pn), f0, f1, reading)
        t = op[1]                                                        # clear
        assert all(pulled[r * W + q] == t for q in range(W)), f"rank {r} clears its tally before everyone pulled it (sweep {t})"
        ta = list(tally); ta[r] = 0
        return (tuple(ta), pulled, P, f0, f1, reading)
    return step


def _p_init(W):
    P = [0] * (W * 2 * W)
    for r in range(W):
        for b in range(W):
            P[(r * 2 + 1) * W + b] = 1                                   # sweep 1 reads buffer 1, drawn before the loop (every rank, all blocks)
    return ((0,) * W, (0,) * (W * W), tuple(P), (0,) * (W * W), (0,) * (W * W), (-1,) * W)


def test_tally_to_p_exchange_has_no_hazard():
    for W, sweeps in ((2, 3), (3, 2)):
        n = _explore(W, [_p_program(W, sweeps)] * W, _p_init(W), _p_step(W))
        assert n > 100


def test_clearing_the_tally_before_the_second_signal_would_be_unsafe():
    """Moving `clear` in front of signal 1 (i.e. clearing right after the own pull) lets a rank wipe a tally a slower rank
    still has to pull."""
    W, sweeps = 2, 2
    prog = []
    for t in range(1, sweeps + 1):
        prog += [("zq_begin", t), ("zq_end", t)]
        prog += [("f0", p, t) for p in range(W)] + [("w0", t)]
        prog += [("pull", q, t) for q in range(W)]
        prog += [("clear", t)]
        prog += [("push", p, t) for p in range(W)]
        prog += [("f1", p, t) for p in range(W)] + [("w1", t)]
    try:
        _explore(W, [prog] * W, _p_init(W), _p_step(W))
    except AssertionError as e:
        assert "clears its tally" in str(e) or "pulls tally" in str(e)
        return
    raise AssertionError("expected a hazard")
