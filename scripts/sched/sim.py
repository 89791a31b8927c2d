"""Development aid: event simulation of per-warp elimination programs for the K4 evaluation kernel.

A program is a list of ops per warp:
  ("visit", j, nslots, wait)   apply column j to nslots accumulator slots (wait: needs ready[j])
  ("publish", i)               row i final -> ready[i]
  ("handoff_arrive", key) / ("handoff_wait", key)
  ("var", j)                   variance tile product of row j (needs ready[j])
Cost model (cycles): FFMA2 issue slots on a scheduler shared by warps w and w+4.
"""
import heapq, sys

SOLO = 0.30      # FFMA2 per cycle for a warp alone on its scheduler inside the FMA region
BOTH = 0.40      # total when both warps are inside it
VISIT_OVH = 350  # non-FMA cycles per visit (control, staging waits)
PUB_LAT = 250    # store + arrive + wake-up latency
FFMA2_PER_SLOT = 512   # one 32x32x32 tile product on a 8-column lane tile

def simulate(progs, nb, verbose=False, detail=False):
    nw = len(progs)
    pc = [0] * nw
    t = [0.0] * nw                 # local clock of each warp (time its current op may start)
    ready = {}                     # barrier -> time
    remaining = [0.0] * nw         # FFMA2 left in the current FMA burst
    state = ["idle"] * nw          # idle | fma | blocked
    block_key = [None] * nw
    now = 0.0
    busy = [0.0] * 4
    # simple time-stepped processor sharing (events are coarse, a step of 50 cycles is plenty)
    DT = 50.0
    fma_need = [0.0] * nw
    ovh_until = [0.0] * nw
    done = [False] * nw
    fin = [0.0] * nw
    waits = [0.0] * nw
    while not all(done):
        # advance control of each warp
        for w in range(nw):
            while not done[w] and state[w] != "fma" and ovh_until[w] <= now:
                if pc[w] >= len(progs[w]):
                    done[w] = True; fin[w] = now; break
                op = progs[w][pc[w]]
                k = op[0]
                if k == "visit":
                    _, j, ns, wait = op
                    if wait and ready.get(("r", j), 1e18) > now:
                        state[w] = "blocked"; break
                    state[w] = "fma"; fma_need[w] = ns * FFMA2_PER_SLOT; pc[w] += 1
                elif k == "var":
                    _, j = op
                    if ready.get(("r", j), 1e18) > now:
                        state[w] = "blocked"; break
                    state[w] = "fma"; fma_need[w] = 1.5 * FFMA2_PER_SLOT; pc[w] += 1   # 4x8 lane tile: smem-bound
                elif k == "publish":
                    ready[("r", op[1])] = now + PUB_LAT; pc[w] += 1; ovh_until[w] = now + 100
                elif k == "handoff_arrive":
                    ready[("h", op[1])] = now + PUB_LAT; pc[w] += 1; ovh_until[w] = now + 300
                elif k == "handoff_wait":
                    if ready.get(("h", op[1]), 1e18) > now:
                        state[w] = "blocked"; break
                    pc[w] += 1; ovh_until[w] = now + 300
                elif k == "ovh":
                    pc[w] += 1; ovh_until[w] = now + op[1]
                else:
                    raise ValueError(k)
                if state[w] == "blocked": state[w] = "idle"
        for w in range(nw):
            if state[w] == "blocked": waits[w] += DT; state[w] = "idle"
        # FMA progress
        for s in range(4):
            ws = [w for w in (s, s + 4) if w < nw and state[w] == "fma"]
            if not ws: continue
            rate = SOLO if len(ws) == 1 else BOTH / 2
            for w in ws:
                fma_need[w] -= rate * DT
                busy[s] += rate * DT
                if fma_need[w] <= 0:
                    state[w] = "idle"; ovh_until[w] = now + DT + VISIT_OVH
        now += DT
        if now > 5e7: raise RuntimeError("deadlock / runaway: pcs %s" % pc)
    total_ffma2 = sum(busy)
    util = total_ffma2 / (4 * 0.5 * now)
    if detail: return now, util, waits, fin, ready
    return now, util, waits

# ---------------------------------------------------------------- schedules
def sched_current(nb, W=8, R=4):
    """The round-1 kernel: waves of W*R rows (partial first), snake dealing, lookahead publish, variance by owner."""
    wave = W * R
    first = nb % wave or wave
    starts = [0]
    while starts[-1] + (first if len(starts) == 1 else wave) < nb:
        starts.append(starts[-1] + (first if len(starts) == 1 else wave))
    progs = [[] for _ in range(W)]
    for w in range(W):
        p = progs[w]
        p.append(("ovh", 28000))     # k* + mean
        if w == 0: p.append(("publish", 0))
        for ci, b in enumerate(starts):
            e = min(nb, b + (first if ci == 0 else wave))
            m = e - b
            rows = [b + (r * W + (W - 1 - w if r & 1 else w)) for r in range(R) if (r * W + (W - 1 - w if r & 1 else w)) < m]
            if not rows: continue
            for j in range(0, rows[-1]):
                act = [i for i in rows if i > j]
                if not act: break
                wait = True
                if act[0] == j + 1:
                    p.append(("visit", j, 1, wait)); p.append(("publish", j + 1))
                    if len(act) > 1: p.append(("visit", j, len(act) - 1, False))
                else:
                    p.append(("visit", j, len(act), wait))
        for ci, b in enumerate(starts):
            e = min(nb, b + (first if ci == 0 else wave))
            m = e - b
            for r in range(R):
                o = r * W + (W - 1 - w if r & 1 else w)
                if o < m: p.append(("var", b + o))
    return progs



def sched_pairs(nb, dfrac=None, wave=16, R=4, var_mode="greedy", kstar=28000):
    """Two teams of 4 warps; waves of 16 rows; the warps tw and tw+4 hold the same rows of every wave.
    Team c&1 is on duty for wave c (streaming columns, triangle, publishing); the other team helps with a
    suffix of the old bulk columns after finishing its own previous triangle."""
    first = nb % wave or wave
    starts = [0]
    while True:
        nxt = starts[-1] + (first if len(starts) == 1 else wave)
        if nxt >= nb: break
        starts.append(nxt)
    ends = [min(nb, s + (first if i == 0 else wave)) for i, s in enumerate(starts)]
    progs = [[] for _ in range(8)]
    for w in range(8): progs[w].append(("ovh", kstar))
    progs[0].append(("publish", 0))
    for c, b in enumerate(starts):
        e = ends[c]; m = e - b
        bp = starts[c - 1] if c > 0 else 0
        D = c & 1
        # duty share of the old bulk [0, bp): [0, h)
        if c == 0 or bp == 0: h = bp
        else:
            d = dfrac(nb, c, bp, b, e) if callable(dfrac) else (dfrac if dfrac is not None else 0.5 + 0.75 * 0.5 * (m / 2.0) / max(bp, 1))
            h = max(0, min(bp, int(round(d * bp))))
        for tw in range(4):
            offs = [r * 4 + (3 - tw if r & 1 else tw) for r in range(R)]
            rows = [b + o for o in offs if o < m]
            if not rows: continue
            duty = progs[tw + 4 * D]; helper = progs[tw + 4 * (1 - D)]
            if h < bp:
                for j in range(h, bp): helper.append(("visit", j, len(rows), True))
                helper.append(("handoff_arrive", (c, tw)))
            for j in list(range(0, h)) + list(range(bp, b)):
                duty.append(("visit", j, len(rows), True))
            if h < bp: duty.append(("handoff_wait", (c, tw)))
            if rows[0] == b and b > 0: duty.append(("publish", b))
            for j in range(b, rows[-1]):
                act = [i for i in rows if i > j]
                if not act: break
                if act[0] == j + 1:
                    duty.append(("visit", j, 1, True)); duty.append(("publish", j + 1))
                    if len(act) > 1: duty.append(("visit", j, len(act) - 1, False))
                else:
                    duty.append(("visit", j, len(act), True))
    # variance rows: greedy on the simulated finish times
    if var_mode == "greedy":
        T, u, waits, fin, rdy = simulate(progs, nb, detail=True)
        avail = list(fin)
        cost = 1.5 * FFMA2_PER_SLOT / SOLO + VISIT_OVH
        order = sorted(range(nb), key=lambda j: rdy.get(("r", j), 0))
        assign = [[] for _ in range(8)]
        for j in order:
            w = min(range(8), key=lambda w: max(avail[w], rdy.get(("r", j), 0)))
            avail[w] = max(avail[w], rdy.get(("r", j), 0)) + cost
            assign[w].append(j)
        for w in range(8):
            for j in assign[w]: progs[w].append(("var", j))
    return progs


if __name__ == "__main__":
    for nb in (20, 30, 36, 41, 48, 53, 64, 80):
        T, u, waits = simulate(sched_current(nb), nb)
        best = None
        for d in (0.5, 0.6, 0.7, 0.8, 0.9, 1.0):
            T2, u2, w2 = simulate(sched_pairs(nb, d), nb)
            if best is None or T2 < best[0]: best = (T2, u2, d, w2)
        print(f"nb={nb:3d} current {T/1e3:7.1f} kcyc util {u:.3f} | pairs {best[0]/1e3:7.1f} kcyc util {best[1]:.3f} d={best[2]} speedup {T/best[0]:.2f} waits {[int(x/1e3) for x in best[3]]}")
