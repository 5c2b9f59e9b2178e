"""CPU tier: tensor-memory hazard check of the fused MLP program (Builder::finalize in csrc/pnr_api.cu).

The kernel keeps accumulators, activations (16-bit hi / lo parts) and head activations in overlapping tensor-memory
column ranges and orders the MMA stages against the two half-epilogues of every step with five barriers whose
placement (F_WAIT_E0 / F_WAIT_E1 / F_COMMIT_ACC0 / F_COMMIT_ACC1 / F_COMMIT_WAR) is computed on the host.
This test rebuilds the happens-before graph those flags imply, over two consecutive tiles, and requires that every
pair of events that touch overlapping columns with at least one write is ordered by it."""
import itertools

import pytest

from panopticnerf_b200 import make_cfg, make_network, synthetic as S
from test_cpu_program import (A_TMEM, EPI_LINEAR_TO_A, EPI_RELU_TO_A, F_COMMIT_ACC0, F_COMMIT_ACC1, F_COMMIT_WAR,
                              F_WAIT_E0, F_WAIT_E1, build)

F_COMMIT_WAR1 = 1024   # split-war programs (PNR_PROGRAM_SPLIT_WAR): second write-after-read barrier of a step


def overlap(a, b):
    return a[0] < a[1] and b[0] < b[1] and a[0] < b[1] and b[0] < a[1]


def events_and_edges(prog, tiles=2):
    """Events: ('S', t, i) MMA stage i of tile t; ('L0', t, s) loads of E0 of step s; ('W0', t, s) / ('W0b', t, s) its
    stores to the lower / upper half of the columns it overwrites (one barrier each in split-war programs, the same
    barrier otherwise); ('E1', t, s) loads + stores of E1.  Returns (reads, writes, edges)."""
    x3 = prog.passes == 3
    split_war = any(prog.st[i].flags & F_COMMIT_WAR1 for i in range(prog.n_stages))
    steps = []
    for i in range(prog.n_stages):
        if prog.st[i].flags & F_WAIT_E0:
            steps.append([])
        steps[-1].append(i)
    reads, writes, edges = {}, {}, []
    order = []            # (tile, step) in execution order
    for t in range(tiles):
        for s in range(len(steps)):
            order.append((t, s))
    prev_stage = None
    for t, s in order:
        ed = prog.ep[s]
        to_a = ed.kind in (EPI_RELU_TO_A, EPI_LINEAR_TO_A)
        for i in steps[s]:
            sd = prog.st[i]
            ev = ("S", t, i)
            r = [(sd.acc_col, sd.acc_col + sd.n)]
            if sd.a_kind == A_TMEM:
                r.append((sd.a_off, sd.a_off + 8 * sd.ksteps))
                if x3:
                    r.append((sd.a_lo_off, sd.a_lo_off + 8 * sd.ksteps))
            reads[ev], writes[ev] = r, [(sd.acc_col, sd.acc_col + sd.n)]
            if prev_stage is not None:
                edges.append((prev_stage, ev))        # the tensor pipe retires MMAs in issue order
            prev_stage = ev
            if sd.flags & F_COMMIT_ACC0:
                edges.append((ev, ("L0", t, s)))
            if sd.flags & F_COMMIT_ACC1:
                edges.append((ev, ("E1", t, s)))
            if sd.flags & F_COMMIT_WAR:
                edges.append((ev, ("W0", t, s)))
                if not split_war:
                    edges.append((ev, ("W0b", t, s)))
            if sd.flags & F_COMMIT_WAR1:
                edges.append((ev, ("W0b", t, s)))
        split = ed.n0 < ed.n
        # a step issued as one half signals only acc_full[1]; its E0 then has all the columns and starts with E1
        if not any(prog.st[i].flags & F_COMMIT_ACC0 for i in steps[s]):
            last = ("S", t, steps[s][-1])
            edges.append((last, ("L0", t, s)))
        reads[("L0", t, s)], writes[("L0", t, s)] = [(ed.acc_col, ed.acc_col + ed.n0)], []
        g0 = ed.n0 // 16
        n0a = (g0 // 2) * 16 if g0 // 2 > 0 else ed.n0          # the kernel's first E0 block (same formula)
        w0, w0b = [], []
        if to_a:
            w0 = [(ed.dst_col, ed.dst_col + n0a // 2)] + ([(ed.dst_lo_col, ed.dst_lo_col + n0a // 2)] if x3 else [])
            w0b = [(ed.dst_col + n0a // 2, ed.dst_col + ed.n0 // 2)] + (
                [(ed.dst_lo_col + n0a // 2, ed.dst_lo_col + ed.n0 // 2)] if x3 else [])
        reads[("W0", t, s)], writes[("W0", t, s)] = [], w0
        reads[("W0b", t, s)], writes[("W0b", t, s)] = [], w0b
        r1 = [(ed.acc_col + ed.n0, ed.acc_col + ed.n)] if split else []
        w1 = []
        if to_a and split:
            w1 = [(ed.dst_col + ed.n0 // 2, ed.dst_col + ed.n // 2)] + (
                [(ed.dst_lo_col + ed.n0 // 2, ed.dst_lo_col + ed.n // 2)] if x3 else [])
        reads[("E1", t, s)], writes[("E1", t, s)] = r1, w1
        edges += [(("L0", t, s), ("W0", t, s)), (("W0", t, s), ("W0b", t, s)), (("W0b", t, s), ("E1", t, s))]  # same warps
    for (t, s), (t2, s2) in zip(order, order[1:]):
        edges.append((("E1", t, s), ("L0", t2, s2)))
        for i in steps[s2]:
            sd = prog.st[i]
            if sd.flags & F_WAIT_E0:
                edges.append((("W0b", t, s), ("S", t2, i)))
            if sd.flags & F_WAIT_E1:
                edges.append((("E1", t, s), ("S", t2, i)))
    return reads, writes, edges


def reachability(nodes, edges):
    idx = {n: k for k, n in enumerate(nodes)}
    succ = [[] for _ in nodes]
    for a, b in edges:
        succ[idx[a]].append(idx[b])
    reach = []
    for k in range(len(nodes)):
        seen, stack = set(), [k]
        while stack:
            v = stack.pop()
            for w in succ[v]:
                if w not in seen:
                    seen.add(w)
                    stack.append(w)
        reach.append(seen)
    return idx, reach


@pytest.mark.parametrize("preset,over", [
    ("cfg1", {}), ("cfg2", {}), ("cfg3", {}), ("cfg2", dict(precision="fp16")),
    ("cfg2", dict(D=5, W=128, num_classes=7, num_instances=3)), ("cfg3", dict(precision="bf16x3", W=128)),
    ("cfg1", dict(D=3, num_classes=45))])
def test_every_tensor_memory_conflict_is_ordered(preset, over):
    cfg = make_cfg(preset, **over)
    prog, _, _ = build(cfg, S.init_network_weights(make_network(cfg), seed=0))
    reads, writes, edges = events_and_edges(prog)
    nodes = list(reads)
    idx, reach = reachability(nodes, edges)
    checked = 0
    for a, b in itertools.combinations(nodes, 2):
        if a[0] == "S" and b[0] == "S":
            continue                                    # MMAs are ordered among themselves by construction
        conflict = any(overlap(x, y) for x in writes[a] for y in reads[b] + writes[b]) or \
                   any(overlap(x, y) for x in writes[b] for y in reads[a])
        if not conflict:
            continue
        checked += 1
        assert idx[b] in reach[idx[a]] or idx[a] in reach[idx[b]], \
            f"{preset} {over}: {a} and {b} touch overlapping tensor-memory columns but are not ordered"
    assert checked > 20


@pytest.mark.parametrize("preset,over", [("cfg1", {}), ("cfg2", {}), ("cfg3", {}), ("cfg2", dict(precision="fp16")),
                                         ("cfg2", dict(D=5, W=128, num_classes=7, num_instances=3))])
def test_split_war_programs_are_ordered_too(preset, over):
    """PNR_PROGRAM_SPLIT_WAR (staged kernel variant): E0's stores are released in two blocks; the lower block's
    barrier must come no later than the single barrier of the product program."""
    cfg = make_cfg(preset, **over)
    net = S.init_network_weights(make_network(cfg), seed=0)
    prog, _, _ = build(cfg, net, flags=2)
    base, _, _ = build(cfg, net)
    reads, writes, edges = events_and_edges(prog)
    nodes = list(reads)
    idx, reach = reachability(nodes, edges)
    for a, b in itertools.combinations(nodes, 2):
        if a[0] == "S" and b[0] == "S":
            continue
        conflict = any(overlap(x, y) for x in writes[a] for y in reads[b] + writes[b]) or \
                   any(overlap(x, y) for x in writes[b] for y in reads[a])
        if conflict:
            assert idx[b] in reach[idx[a]] or idx[a] in reach[idx[b]], f"{preset} {over}: {a} / {b} unordered"
    war = [i for i in range(prog.n_stages) if prog.st[i].flags & F_COMMIT_WAR]
    war1 = [i for i in range(prog.n_stages) if prog.st[i].flags & F_COMMIT_WAR1]
    war_base = [i for i in range(base.n_stages) if base.st[i].flags & F_COMMIT_WAR]
    assert len(war) == len(war1) == len(war_base) == prog.n_steps
    assert all(a <= b for a, b in zip(war, war_base))
    if cfg.W >= 256 and cfg.precision.endswith("x3"):    # otherwise one weight stage (K = 64 / 128) covers a whole block
        assert any(a < b for a, b in zip(war, war_base))
    assert war1 == war_base                              # the upper block is released where the single barrier was


def test_the_checker_sees_a_missing_wait():
    """Remove one F_WAIT_E1 and the same analysis must find an unordered conflict (the check is not vacuous)."""
    cfg = make_cfg("cfg2")
    prog, _, _ = build(cfg, S.init_network_weights(make_network(cfg), seed=0))
    victim = next(i for i in range(prog.n_stages) if prog.st[i].flags & F_WAIT_E1 and not prog.st[i].flags & F_WAIT_E0)
    prog.st[victim].flags &= ~F_WAIT_E1
    reads, writes, edges = events_and_edges(prog)
    nodes = list(reads)
    idx, reach = reachability(nodes, edges)
    bad = 0
    for a, b in itertools.combinations(nodes, 2):
        if a[0] == "S" and b[0] == "S":
            continue
        conflict = any(overlap(x, y) for x in writes[a] for y in reads[b] + writes[b]) or \
                   any(overlap(x, y) for x in writes[b] for y in reads[a])
        if conflict and not (idx[b] in reach[idx[a]] or idx[a] in reach[idx[b]]):
            bad += 1
    assert bad > 0
