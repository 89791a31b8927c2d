"""The elimination programs that drive k_eval_v3 (gpismap_b200/csrc/query_v3.cuh) are generated on the host; a
wrong program would hang or corrupt the solve on the GPU, so their invariants are checked here on the CPU:
coverage (every block row meets every earlier column exactly once), tile indices, the publish protocol,
freedom from deadlock under the kernel's wait rules, and the variance rows forming a partition."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpismap_b200", "libgpis_b200.so")
WARPS = 8


def _tile_index(bi, bk, nb):
    return bk * nb - bk * (bk - 1) // 2 + (bi - bk)


def _program(lib, nb, warp):
    n = lib.gpis_debug_program(nb, warp, None, 0)
    assert n > 0
    buf = np.zeros(n, np.int32)
    assert lib.gpis_debug_program(nb, warp, buf.ctypes.data_as(C.c_void_p), n) == n
    r = buf.reshape(-1, 4)
    nwaves, nvis, nvar = int(r[0, 0]), int(r[0, 1]), int(r[0, 2])
    waves = [(int(r[2 + 2 * c, 0]), [int(v) for v in r[3 + 2 * c]]) for c in range(nwaves)]
    v0 = 2 + 2 * nwaves
    visits = []

    def tile16(word, hi):
        t = (int(word) >> 16) & 0xFFFF if hi else int(word) & 0xFFFF
        return -1 if t == 0xFFFF else t

    for v in range(nvis):
        a = r[v0 + v]
        f = int(a[0]) & 0xFFFFFFFF
        tiles = [tile16(a[1], 0), tile16(a[1], 1), tile16(a[2], 0), tile16(a[2], 1)]
        visits.append(dict(j=(f >> 24) & 255, s_lo=f & 7, cnt=(f >> 3) & 7, part=(f >> 6) & 1, split=(f >> 7) & 1,
                           valid=(f >> 8) & 1, wave=(f >> 16) & 255, solo_tile=tiles[f & 7], tiles=tiles))
    for t in range(4):   # terminators
        assert ((int(r[v0 + nvis + t, 0]) >> 8) & 1) == 0
    var_rows = [int(x) for x in r[v0 + nvis + 4:].ravel()[:nvar]]
    return waves, visits, var_rows


@pytest.mark.parametrize("nb", [1, 2, 7, 8, 9, 31, 32, 33, 40, 41, 47, 53, 64, 65, 80])
def test_program_invariants(nb):
    if not os.path.exists(LIB):
        pytest.skip("libgpis_b200.so not built")
    lib = C.CDLL(LIB)
    lib.gpis_debug_program.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
    progs = [_program(lib, nb, w) for w in range(WARPS)]
    covered = np.zeros((nb, nb), np.int32)
    owner = {}
    for w, (waves, visits, var_rows) in enumerate(progs):
        for R, rows in waves:
            assert 0 <= R <= 4
            act = rows[:R]
            assert act == sorted(act, reverse=True)           # slot s holds the (R-1-s)-th row, ascending
            for i in act:
                assert i not in owner
                owner[i] = w
        for v in visits:
            assert v["valid"] == 1
            R, rows = waves[v["wave"]]
            for s in range(4):
                if v["s_lo"] <= s < v["s_lo"] + v["cnt"]:
                    i = rows[s]
                    assert i > v["j"]
                    assert v["tiles"][s] == _tile_index(i, v["j"], nb)
                    covered[i, v["j"]] += 1
                else:
                    assert v["tiles"][s] == -1
            assert v["solo_tile"] == v["tiles"][v["s_lo"]]
            if v["part"] == 0 and v["split"]:
                assert v["cnt"] == 1 and rows[v["s_lo"]] == v["j"] + 1   # lookahead: the row that becomes final
    assert sorted(owner) == list(range(nb))
    want = np.tril(np.ones((nb, nb), np.int32), -1)
    assert np.array_equal(covered, want)
    # variance rows: a partition of all block rows
    allvar = sorted(r for _, _, vr in progs for r in vr)
    assert allvar == list(range(nb))
    # dataflow execution: part-0 visits wait for ready[j]; the solo part publishes row j+1; nothing may block forever
    ready = {0}
    pcs = [0] * WARPS
    done_cols = [dict() for _ in range(WARPS)]   # row -> set of applied columns (rows are final when all j < i are in)
    progress = True
    while progress:
        progress = False
        for w, (waves, visits, _) in enumerate(progs):
            while pcs[w] < len(visits):
                v = visits[pcs[w]]
                if v["part"] == 0 and v["j"] not in ready:
                    break
                R, rows = waves[v["wave"]]
                for s in range(v["s_lo"], v["s_lo"] + v["cnt"]):
                    done_cols[w].setdefault(rows[s], set()).add(v["j"])
                if v["part"] == 0 and v["split"]:
                    i = v["j"] + 1
                    assert done_cols[w][i] == set(range(i)), "row published before all its columns were applied"
                    ready.add(i)
                pcs[w] += 1
                progress = True
    assert all(pcs[w] == len(progs[w][1]) for w in range(WARPS)), "deadlock in the elimination programs"
    assert ready == set(range(nb))
    # every variance row is published by somebody (the kernel waits on ready[j] before using U_j)
    assert set(allvar) <= ready
