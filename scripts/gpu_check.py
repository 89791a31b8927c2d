"""Diagnostic run on a GPU box: prints parity numbers for K1/K2/K4 against the oracle and the
reference build (not a test; tests/ holds the asserting versions)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from gpismap_b200 import cabi  # noqa: E402
from oracle import oraclepy, refpy  # noqa: E402

small = "--small" in sys.argv
O = oraclepy.Oracle()
rng = np.random.default_rng(0)


def relmax(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def leaf_parity(dim):
    P = H.P3 if dim == 3 else H.P2
    ctx = cabi.Ctx(dim)
    sizes = [1, 2, 5, 31, 40, 100] + ([] if small else [200, 260, 330])
    cells, centres, offs, chunks = [], [], [0], []
    for i, N in enumerate(sizes):
        s = H.leaf_samples3(N, rng) if dim == 3 else H.leaf_samples2(N, rng)
        chunks.append(s)
        offs.append(offs[-1] + N)
        cell = [i, 0, 0][:dim]
        cells.append(cell)
        centres.append([(2 * c + 1) * P["half"] for c in cell])
    st = ctx.leaves_update(cells, centres, offs, np.concatenate(chunks))
    print(f"[leaf dim={dim}] status {st.tolist()} train_ms {ctx.stats()['last_train_ms']:.3f}")
    for i, N in enumerate(sizes):
        g = O.gp_train(dim, chunks[i], P["scale"], P["noise"])
        a, L, gf = g.factors()
        got = ctx.leaf_get(cells[i])
        assert got["n"] == g.n, (got["n"], g.n)
        print(f"  N={N:4d} n={g.n:4d} alpha rel {relmax(got['alpha'], a):.2e}  L rel {relmax(got['L'], L):.2e} "
              f"gf ok {bool((got['gradflag'] == gf).all())} nan {int(np.isnan(got['alpha']).sum())}")
    ctx.close()


def obs_parity():
    ctx = cabi.Ctx(3)
    ni, nj = (40, 56) if small else (240, 320)
    v = (np.arange(ni) * 2 - 224) / 568.0
    u = (np.arange(nj) * 2 - 310) / 568.0
    vu = np.zeros((nj, ni, 2), np.float32)
    vu[:, :, 0] = v[None, :]
    vu[:, :, 1] = u[:, None]
    z = 1.5 + 0.3 * np.sin(3 * vu[:, :, 0]) + 0.2 * np.cos(2 * vu[:, :, 1])
    zinv = (1.0 / z).astype(np.float32)
    zinv[rng.uniform(size=zinv.shape) < 0.1] = -1.0
    zinv[:12, :9] = -1.0
    t0 = time.time()
    ctx.obs_train_2d(vu, zinv, ni, nj)
    t1 = time.time()
    o = O.obs2d(vu, zinv, ni, nj)
    m = 4000
    xt = np.stack([rng.uniform(v[0] - 0.01, v[-1] + 0.01, m), rng.uniform(u[0] - 0.01, u[-1] + 0.01, m)], 1).astype(np.float32)
    val0 = rng.uniform(size=m).astype(np.float32)
    t2 = time.time()
    gv, gr = ctx.obs_test(xt, 2, val=val0)
    t3 = time.time()
    ov, orr = o.test(xt, val=val0)
    ev = orr < 1e5
    print(f"[obs2d] train {1e3*(t1-t0):.1f} ms test {1e3*(t3-t2):.1f} ms; evaluated {int(ev.sum())}/{m}; "
          f"mask eq {bool(((gr < 1e5) == ev).all())}; val rel {relmax(gv[ev], ov[ev]):.2e} var abs {np.abs(gr[ev]-orr[ev]).max():.2e} "
          f"untouched ok {bool((gv[~ev] == val0[~ev]).all())}")
    # 1-D
    ctx2 = cabi.Ctx(2)
    th = np.linspace(-2.35, 2.35, 270).astype(np.float32)
    f = (1.0 / np.sqrt(2.0 + np.sin(2 * th))).astype(np.float32)
    ctx2.obs_train_1d(th, f)
    o1 = O.obs1d(th, f)
    xt1 = rng.uniform(-2.4, 2.4, 3000).astype(np.float32)
    gv, gr = ctx2.obs_test(xt1, 1)
    ov, orr = o1.test(xt1)
    ev = orr < 1e5
    print(f"[obs1d] evaluated {int(ev.sum())}; mask eq {bool(((gr < 1e5) == ev).all())}; val rel {relmax(gv[ev], ov[ev]):.2e} "
          f"var abs {np.abs(gr[ev]-orr[ev]).max():.2e}")
    ctx.close(); ctx2.close()


def map_parity(dim):
    P = H.P3 if dim == 3 else H.P2
    if dim == 3:
        M = refpy.RefMap3()
        s = H.sphere_samples(0.21 if small else 0.33, 0.0125, (0.0517, 0.0231, 0.0113), rng)
    else:
        M = refpy.RefMap2()
        s = H.circle_samples(6.3, 0.45, (1.7, -2.3), rng)
    ins = M.insert_samples(s)
    t0 = time.time()
    nact = M.update_gps()
    t1 = time.time()
    centres, offs, samples, trained = H.ref_map_to_csr(M, P)
    Ns = np.diff(offs)
    print(f"[map dim={dim}] inserted {ins}/{len(s)} clusters {len(centres)} active {nact} N p50/max {int(np.median(Ns))}/{Ns.max()} ref train {t1-t0:.2f}s")
    ctx = cabi.Ctx(dim)
    rm, lv = H.root_cells(M, P)
    ctx.rebase(rm, lv)
    cells = cabi.cells_of(centres, P["half"])
    st = ctx.leaves_update(cells, centres, offs, samples)
    stt = ctx.stats()
    print(f"  gpu train {stt['last_train_ms']:.2f} ms  ({stt['last_train_flops']/stt['last_train_ms']/1e9:.2f} TFLOP/s) bad pivots {int((st>0).sum())}")
    # queries: shell around the surface + far points
    nq = 3000 if small else 20000
    if dim == 3:
        d = rng.normal(size=(nq, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rad = (0.21 if small else 0.33) + rng.uniform(-0.12, 0.12, nq)
        x = (d * rad[:, None] + np.array([0.0517, 0.0231, 0.0113])).astype(np.float32)
    else:
        d = rng.normal(size=(nq, 2)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rad = 6.3 + rng.uniform(-6.0, 6.0, nq)
        x = (d * rad[:, None] + np.array([1.7, -2.3])).astype(np.float32)
    w = 2 * (1 + dim)
    init = rng.uniform(size=(nq, w)).astype(np.float32)
    t0 = time.time(); want = M.test(x, init.copy()); t1 = time.time()
    gps = [O.gp_train(dim, samples[offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(centres))]
    om = O.make_map(dim, centres, P["half"], gps, P["search"], P["var_thre"], P["noise"])
    ores, ochosen, otie = om.test(x, init.copy(), want_choice=True)
    print(f"  ref test {nq/(t1-t0):.0f} q/s; oracle vs ref max abs {np.abs(ores-want).max():.2e}; ties {int(otie.sum())}")
    slot_of = np.array([ctx.leaf_index(c) for c in cells])
    inv = -np.ones(slot_of.max() + 2, np.int64); inv[slot_of] = np.arange(len(slot_of))
    for ver in (1, 2, 3):
        ctx.set_eval_version(ver)
        t0 = time.time(); got, chosen, tie = ctx.query(x, init.copy(), debug=True); t1 = time.time()
        stq = ctx.stats()
        ch = chosen.copy()
        for k in (1, 2, 3):
            ch[:, k] = np.where(chosen[:, k] >= 0, inv[np.maximum(chosen[:, k], 0)], -1)
        ok = (ch == ochosen).all(1) | (otie > 0)
        wv = 1 + dim
        ef = np.abs(got[:, 0] - want[:, 0]) / np.maximum(np.abs(want[:, 0]), 1e-3)
        eg = np.linalg.norm(got[:, 1:wv] - want[:, 1:wv], axis=1) / np.maximum(np.linalg.norm(want[:, 1:wv], axis=1), 1e-2)
        evv = np.abs(got[:, wv:] - want[:, wv:]) / np.maximum(np.abs(want[:, wv:]), 1e-3)
        print(f"  v{ver}: neighbour sets equal {int(ok.sum())}/{nq} tie flags eq {bool((tie==otie).all())}; f {np.nanmax(ef):.2e} grad {np.nanmax(eg):.2e} var {np.nanmax(evv):.2e} "
              f"nan {int(np.isnan(got).sum())}; evals {stq['last_query_evals']} kernel {stq['last_query_ms']:.2f} ms (eval {stq['last_query_eval_ms']:.2f}) wall {1e3*(t1-t0):.1f} ms")
        bad = np.where(~(ef < 1e-4) | ~(eg < 1e-4) | ~(evv.max(1) < 1e-3))[0]
        if len(bad):
            b = bad[0]
            print("   first bad", b, "nc", chosen[b], "\n   got ", got[b], "\n   want", want[b], "\n   orcl", ores[b])
    ctx.close()


def train_throughput():
    ctx = cabi.Ctx(3)
    nl = 64 if small else 1184
    N = 100 if small else 200
    cells, centres, offs, chunks = [], [], [0], []
    for i in range(nl):
        s = H.leaf_samples3(N, rng)
        s[:, 8] = 0.02
        chunks.append(s); offs.append(offs[-1] + N)
        cell = [i % 64, i // 64, 0]
        cells.append(cell); centres.append([(2 * c + 1) * 0.025 for c in cell])
    samples = np.concatenate(chunks)
    for rep in range(3):
        st = ctx.leaves_update(cells, centres, offs, samples)
        s = ctx.stats()
        print(f"[train x{nl}] n={s['last_train_sum_n']//nl} {s['last_train_ms']:.2f} ms -> {s['last_train_flops']/s['last_train_ms']/1e9:.2f} TFLOP/s, "
              f"{nl/s['last_train_ms']*1e3:.0f} leaves/s; bad {int((st>0).sum())}")
    ctx.close()


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["leaf", "obs", "map", "thr"]
    if "leaf" in which:
        leaf_parity(3); leaf_parity(2)
    if "obs" in which:
        obs_parity()
    if "map" in which:
        map_parity(3); map_parity(2)
    if "thr" in which:
        train_throughput()
