"""A small event model of one CTA of the fused MLP kernel: the tensor pipe (serial MMA stages), the epilogue warps
(in order E0, E1 of every step) and the barriers between them, driven by the REAL per-tile program
(`pnr_program_host`, CPU only).  Calibrated against the clock64 timeline of the product kernel
(profiles/r01_timeline_v10_fp16x3.log: ~71.8 k cycles per cfg2 tile, trunk layer ~8000 cycles, E0 hand-off ~580,
E1 hand-off ~1000); used to rank schedule changes before spending GPU time on them.

    python tools/schedule_model.py [preset] [precision]

It is a planning aid, not a measurement: everything it prints is labelled "model"."""
import sys
from dataclasses import dataclass
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from panopticnerf_b200 import make_cfg, make_network, synthetic as S      # noqa: E402
import test_cpu_program as T                                              # noqa: E402

F_WAIT_E0, F_WAIT_E1, F_ACC0, F_ACC1, F_WAR, F_WAR1 = 2, 4, 8, 16, 32, 1024


@dataclass
class K:                      # cycles; calibrated on the v10 timeline
    mma_min: float = 50.0     # one MMA cannot be issued faster than this (N = 64 stages are issue-bound)
    stage_gap: float = 110.0  # pipe idle between two stages (hand-off between the two issuer warps, queue ~3 deep)
    commit: float = 120.0     # last MMA of a stage retired -> mbarrier phase complete
    wake: float = 150.0       # barrier complete -> sleeping epilogue warp runs (nanosleep back-off)
    first_ld: float = 270.0   # first tcgen05.ld of an epilogue half
    group: float = 450.0      # one 16-column group per warp (convert, split, store), loads overlapped
    store: float = 100.0      # drain of the last tcgen05.st
    signal: float = 450.0     # epilogue done -> scout -> ready counter -> issuer's first MMA
    warps_per_quarter: int = 2


def simulate(prog, k: K, tiles: int = 4, split_war: bool = False, verbose: bool = False):
    x3 = prog.passes == 3
    per_k16 = 3 if x3 else 1
    steps = []
    for i in range(prog.n_stages):
        if prog.st[i].flags & F_WAIT_E0:
            steps.append([])
        steps[-1].append(i)
    pipe_free = 0.0
    epi_free = 0.0
    e_done = [0.0, 0.0]                    # of the previous step
    spans, layer_periods, waits0, waits1 = [], [], [], []
    for t in range(tiles):
        t_start = None
        for s, idxs in enumerate(steps):
            ed = prog.ep[s]
            to_a = ed.kind in (T.EPI_RELU_TO_A, T.EPI_LINEAR_TO_A)
            acc_full = [None, None]
            war_ready = [None, None]
            step_begin = None
            for i in idxs:
                sd = prog.st[i]
                dep = 0.0
                if sd.flags & F_WAIT_E0:
                    dep = max(dep, e_done[0] + k.signal)
                if sd.flags & F_WAIT_E1:
                    dep = max(dep, e_done[1] + k.signal)
                start = max(pipe_free, dep)
                if dep > pipe_free and (t == tiles - 2):
                    (waits0 if sd.flags & F_WAIT_E0 else waits1).append(dep - pipe_free)
                dur = sd.ksteps * per_k16 * max(sd.n / 2.0, k.mma_min) + k.stage_gap
                end = start + dur
                pipe_free = end
                if step_begin is None:
                    step_begin = start
                if t_start is None:
                    t_start = start
                if sd.flags & F_ACC0:
                    acc_full[0] = end + k.commit
                if sd.flags & F_ACC1:
                    acc_full[1] = end + k.commit
                if sd.flags & F_WAR:
                    war_ready[0] = end + k.commit
                if sd.flags & F_WAR1:
                    war_ready[1] = end + k.commit
            if war_ready[1] is None:
                war_ready[1] = war_ready[0]
            if s in (2,) and t == tiles - 2:
                layer_periods.append(step_begin)
            if s in (3,) and t == tiles - 2:
                layer_periods.append(step_begin)
            g0 = ed.n0 // 16
            g1 = (ed.n - ed.n0) // 16
            w = k.warps_per_quarter
            # E0
            begin0 = max(acc_full[0] + k.wake, epi_free)
            per_warp0 = -(-g0 // w)
            if not to_a or per_warp0 == 0:
                end0 = begin0 + (k.first_ld + per_warp0 * k.group if per_warp0 else 0.0)
            elif split_war and per_warp0 >= 2:
                a = per_warp0 // 2
                tA = max(begin0 + k.first_ld + a * k.group, war_ready[0] + k.wake)
                tB = max(tA + k.first_ld * 0.5 + (per_warp0 - a) * k.group, war_ready[1] + k.wake)
                end0 = tB + k.store
            else:
                stash = min(2, per_warp0)
                tW = max(begin0 + k.first_ld + stash * k.group, war_ready[0] + k.wake)
                end0 = tW + (per_warp0 - stash) * k.group + k.store
            # E1
            begin1 = max(acc_full[1] + k.wake, end0)
            per_warp1 = -(-g1 // w)
            end1 = begin1 + (k.first_ld + per_warp1 * k.group + k.store if per_warp1 else 0.0)
            e_done = [end0, end1]
            epi_free = end1
        spans.append(pipe_free - t_start)
    res = {"tile_span": spans[-2], "e0_wait": sum(waits0) / max(len(waits0), 1),
           "e1_wait": sum(waits1) / max(len(waits1), 1),
           "layer": (layer_periods[1] - layer_periods[0]) if len(layer_periods) == 2 else None}
    if verbose:
        print(res)
    return res


def ideal_cycles(prog):
    per_k16 = 3 if prog.passes == 3 else 1
    return sum(prog.st[i].ksteps * per_k16 * prog.st[i].n / 2.0 for i in range(prog.n_stages))


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
    cfg = make_cfg(preset, precision=precision)
    net = S.init_network_weights(make_network(cfg), seed=0)
    prog, _, _ = T.build(cfg, net)
    prog_sw, _, _ = T.build(cfg, net, flags=2)
    ideal = ideal_cycles(prog)
    print(f"{preset} {precision}: {prog.n_stages} stages, {prog.n_steps} steps, tensor work {ideal:.0f} cycles per tile")
    rows = [("product schedule (calibration target: ~71.8 k measured)", prog, K(), False),
            ("+ split write-after-read barrier (-DPNR_SPLIT_WAR)", prog_sw, K(), True),
            ("+ per-warp arrivals (-DPNR_WARP_ARRIVE, signal -100)", prog, K(signal=350.0), False),
            ("+ both", prog_sw, K(signal=350.0), True),
            ("epilogue 20 % cheaper per group (both above too)", prog_sw, K(signal=350.0, group=360.0), True),
            ("no stage gaps (one CTA-pair instruction stream feeds two SMs)", prog, K(stage_gap=0.0), False),
            ("no signalling latency at all (bound of this schedule)", prog_sw, K(signal=0.0, wake=0.0, commit=0.0), True)]
    for name, p, k, sw in rows:
        r = simulate(p, k, split_war=sw)
        print(f"  model: {name:58s} tile {r['tile_span']:8.0f} cycles  pipe busy {100 * ideal / r['tile_span']:5.1f} %"
              f"  layer {r['layer']:6.0f}  E0 wait {r['e0_wait']:5.0f}  E1 wait {r['e1_wait']:5.0f}")


if __name__ == "__main__":
    main()
