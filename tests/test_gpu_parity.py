"""Parity of the CUDA path (through the C ABI) with the oracle, on a real B200: pytest -m gpu.
Nothing here reads /root/reference: the checker is the C restatement oracle (+ its fp64 shadow) and
the committed golden vectors; oracle/_ref is used only where its prebuilt .so travelled with the repo."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cabi():
    from gpismap_b200 import cabi as c
    c.lib()           # raises if the CUDA library is missing: no fallback
    return c


def _train_batch(cabi, dim, sizes, seed):
    rng = np.random.default_rng(seed)
    P = H.P3 if dim == 3 else H.P2
    ctx = cabi.Ctx(dim)
    cells, centres, offs, chunks = [], [], [0], []
    for i, N in enumerate(sizes):
        s = H.leaf_samples3(N, rng) if dim == 3 else H.leaf_samples2(N, rng)
        chunks.append(s)
        offs.append(offs[-1] + N)
        cell = [i, -2, 5][:dim]
        cells.append(cell)
        centres.append([(2 * c + 1) * P["half"] for c in cell])
    samples = np.concatenate(chunks)
    st = ctx.leaves_update(cells, centres, offs, samples)
    return ctx, P, cells, chunks, st


@pytest.mark.parametrize("dim", [3, 2])
def test_leaf_train_matches_oracle(cabi, oracle, oracle64, dim):
    """K1: gradflag rule, covariance, Cholesky, alpha. Factors are compared through what they must
    satisfy (L L^T = K, K alpha = y in fp64) and against the fp32 oracle's factors."""
    sizes = [1, 2, 3, 31, 32, 33, 64, 100, 200, 290]
    ctx, P, cells, chunks, st = _train_batch(cabi, dim, sizes, 5)
    assert (st == 0).all()
    for cell, s in zip(cells, chunks):
        o = oracle.gp_train(dim, s, P["scale"], P["noise"])
        a32, L32, gf = o.factors()
        a64, L64, _ = oracle64.gp_train(dim, s, P["scale"], P["noise"]).factors()
        got = ctx.leaf_get(cell)
        assert got["n"] == o.n and got["N"] == len(s) and got["ng"] == o.ng
        assert np.array_equal(got["gradflag"], gf)
        scale_L = np.abs(L64).max()
        e_gpu = np.abs(got["L"] - L64).max() / scale_L
        e_ref = np.abs(L32 - L64).max() / scale_L
        assert e_gpu < max(4 * e_ref, 2e-6), (len(s), e_gpu, e_ref)
        # alpha is only as well determined as cond(K) allows in fp32: hold the GPU to the oracle's own error
        ea_gpu = np.abs(got["alpha"] - a64).max() / np.abs(a64).max()
        ea_ref = np.abs(a32 - a64).max() / np.abs(a64).max()
        assert ea_gpu < max(4 * ea_ref, 1e-5), (len(s), ea_gpu, ea_ref)
        sx = np.where(gf > 0, s[:, 2 * dim + 1], 2.0).astype(np.float32)
        K = oracle64.matern_train(dim, s[:, :dim], gf, P["scale"], sx, s[:, 2 * dim + 2])
        Lg = got["L"].astype(np.float64)
        assert np.abs(Lg @ Lg.T - K).max() / np.abs(K).max() < 5e-6
    ctx.close()


def test_leaf_capacity_and_empty(cabi):
    ctx = cabi.Ctx(3)
    # an empty training ball only registers the leaf (GPisMap3.cpp:710)
    st = ctx.leaves_update([[0, 0, 0]], [[0.025, 0.025, 0.025]], [0, 0], np.zeros((0, 9), np.float32))
    assert ctx.leaf_index([0, 0, 0]) >= 0 and ctx.leaf_get([0, 0, 0]) is None
    s = ctx.stats()
    assert s["leaves"] == 1 and s["leaves_trained"] == 0
    # a leaf beyond the kernel's capacity is registered but not trained and flagged; the rest of the batch trains,
    # and nothing is left half-mutated (the reference has no size limit, so this must not poison the frame)
    rng = np.random.default_rng(4)
    big = np.zeros((2000, 9), np.float32)
    ok = H.leaf_samples3(40, rng)
    st = ctx.leaves_update([[1, 0, 0], [2, 0, 0]], [[0.075, 0.025, 0.025], [0.125, 0.025, 0.025]], [0, 2000, 2040],
                           np.concatenate([big, ok]))
    assert st[0] < 0 and st[1] == 0
    assert ctx.leaf_index([1, 0, 0]) >= 0 and ctx.leaf_get([1, 0, 0]) is None
    assert ctx.leaf_get([2, 0, 0])["N"] == 40
    assert ctx.stats()["last_train_skipped"] == 1
    with pytest.raises(RuntimeError):
        ctx.leaves_update([[3, 0, 0]], [[0.175, 0.025, 0.025]], [5, 2], ok)      # decreasing offsets: rejected up front
    assert ctx.leaf_index([3, 0, 0]) == -1                                        # ... and nothing was registered
    ctx.leaves_erase([[0, 0, 0]])
    assert ctx.leaf_index([0, 0, 0]) == -1
    ctx.close()


def _fixture_map(cabi):
    g = dict(np.load(os.path.join(G, "map3d.npz")))
    P = H.P3
    ctx = cabi.Ctx(3)
    pitch = 2.0 * np.float64(np.float32(P["half"]))
    root_min = np.round((g["root_c"].astype(np.float64) - float(g["root_half"])) / pitch).astype(np.int32)
    levels = int(round(np.log2(float(g["root_half"]) / np.float64(np.float32(P["half"])))))
    ctx.rebase(root_min, levels)
    cells = cabi.cells_of(g["centres"], P["half"])
    st = ctx.leaves_update(cells, g["centres"], g["offsets"], g["samples"])
    assert (st == 0).all()
    ctx.leaves_set_boxes(cells, g["boxes"])
    return ctx, g, cells, P


@pytest.mark.parametrize("version", [3, 1])
def test_query_matches_reference_fixture(cabi, oracle, oracle64r, version):
    """K3 + K4 against rows produced by the unmodified reference (tests/golden/map3d.npz): candidate
    counts bit-exact (incl. lattice-touching boxes), picks identical where the fixture has exact ties
    and > 16 candidates (std::sort replay), values within the stated tolerances."""
    ctx, g, cells, P = _fixture_map(cabi)
    ctx.set_eval_version(version)
    got, chosen, tie = ctx.query(g["X"], g["init"].copy(), debug=True)
    assert np.array_equal(chosen[:, 0], g["ncand"])
    offs = g["offsets"]
    gps = [oracle.gp_train(3, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    gps64 = [oracle64r.gp_train(3, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    m = oracle.make_map(3, g["centres"], P["half"], gps, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    m64 = oracle64r.make_map(3, g["centres"], P["half"], gps64, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    want, ochosen, otie = m.test(g["X"], g["init"].copy(), want_choice=True)
    want64 = m64.test(g["X"], g["init"].astype(np.float64))
    assert np.array_equal(want, g["rows"])                      # the oracle itself reproduces the fixture
    slot = np.array([ctx.leaf_index(c) for c in cells])
    inv = -np.ones(slot.max() + 2, np.int64)
    inv[slot] = np.arange(len(slot))
    ch = chosen.copy()
    for k in (1, 2, 3):
        ch[:, k] = np.where(chosen[:, k] >= 0, inv[np.maximum(chosen[:, k], 0)], -1)
    assert np.array_equal(ch, ochosen), "neighbour picks differ from the reference's std::sort order"
    assert np.array_equal(tie, otie)
    assert (tie > 0).sum() > 50
    ev = g["ncand"] > 0
    evi = np.flatnonzero(ev)
    H.check_rows(got[ev], g["rows"][ev], want64[ev], 3, label=f"v{version}",
                 explain=lambda i: H.selection_ambiguity(gps64, ochosen[evi[i]], g["X"][evi[i]], P["var_thre"], 3))
    # read-modify-write: rows without candidates keep the caller's contents, var_f preset
    keep = [0, 1, 2, 3, 5, 6, 7]
    assert np.array_equal(got[~ev][:, keep], g["init"][~ev][:, keep])
    assert np.all(got[~ev][:, 4] == np.float32(1.0 + np.float32(5e-3)))
    ctx.close()


def test_larger_training_balls_through_host_class(cabi, oracle, oracle64r):
    """BASELINE configs[4] knob end to end on the GPU: the drop-in GPisMap3 with GPisMap3Tuning.rtimes = 2.5 against
    rows of the reference rebuilt with GPISMAP3_RTIMES 2.5 (tests/golden/map3d_rt25.npz)."""
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "map3d_rt25.npz")))
    P = H.P3
    m = hostapi.GPisMap3(rtimes=2.5)
    assert m.insert_samples(g["samples_in"]) == len(g["samples_in"])
    m.train_active()
    assert np.array_equal(m.leaves()[0], g["centres"])
    got = m.test(g["X"], g["init"].copy())
    offs = g["offsets"]
    gps = [oracle.gp_train(3, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    gps64 = [oracle64r.gp_train(3, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    mo = oracle.make_map(3, g["centres"], P["half"], gps, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    m64 = oracle64r.make_map(3, g["centres"], P["half"], gps64, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    want, ochosen, _ = mo.test(g["X"], g["init"].copy(), want_choice=True)
    assert np.array_equal(want, g["rows"])
    want64 = m64.test(g["X"], g["init"].astype(np.float64))
    ev = g["ncand"] > 0
    evi = np.flatnonzero(ev)
    H.check_rows(got[ev], g["rows"][ev], want64[ev], 3, label="rtimes 2.5",
                 explain=lambda i: H.selection_ambiguity(gps64, ochosen[evi[i]], g["X"][evi[i]], P["var_thre"], 3))
    m.close()


def test_device_gather_equals_host_gather(cabi):
    """f-2: the device-side dirty set + training-set gather (sample store per leaf, flat lookup, DFS-ordered balls)
    must train exactly what the host tree walk trains: two maps of the same frames, one per path, answer a query
    grid bit-identically (identical training sets in identical row order give identical records), in 3-D and 2-D."""
    from gpismap_b200 import hostapi, synth
    X = synth.query_grid(80)
    rows, trained = [], []
    for host_gather in ("1", "0"):
        os.environ["GPIS_HOST_GATHER"] = host_gather
        m = hostapi.GPisMap3()
        n = 0
        for k in range(5):
            dz, pose = synth.frame(k, 40)
            m.update(dz, pose)
            n += m.timing()[1][2]
        rows.append(m.test(X))
        trained.append(int(n))
        m.close()
    os.environ.pop("GPIS_HOST_GATHER")
    assert trained[0] == trained[1] and trained[0] > 2000, trained
    assert (rows[0][:, 4] < 1).sum() > 10000
    assert np.array_equal(rows[0], rows[1])
    g = dict(np.load(os.path.join(G, "seq2d_demo.npz")))
    rows = []
    for host_gather in ("1", "0"):
        os.environ["GPIS_HOST_GATHER"] = host_gather
        m = hostapi.GPisMap()
        for i in range(12):
            m.update(g["thetas"], g["ranges"][i], g["pose6"][i])
        rows.append(m.test(g["X"]))
        m.close()
    os.environ.pop("GPIS_HOST_GATHER")
    assert np.array_equal(rows[0], rows[1])


def test_query_invariances(cabi):
    """Size-independent properties: results do not depend on batch composition (grouping by leaf,
    chunking, order): permuted and split batches are bit-identical to the single batch."""
    ctx, g, cells, P = _fixture_map(cabi)
    rng = np.random.default_rng(0)
    X = np.tile(g["X"][:300], (40, 1)) + rng.normal(0, 0.004, (12000, 3)).astype(np.float32)
    X = X.astype(np.float32)
    base = ctx.query(X)
    perm = rng.permutation(len(X))
    assert np.array_equal(ctx.query(X[perm])[np.argsort(perm)], base)
    parts = np.concatenate([ctx.query(X[:5000]), ctx.query(X[5000:5001]), ctx.query(X[5001:])])
    assert np.array_equal(parts, base)
    # v1 (one CTA per pair) is a different kernel: equal within tolerance only
    ctx.set_eval_version(1)
    v1 = ctx.query(X)
    ev = base[:, 4] < 1.0
    assert np.abs(v1[ev, 0] - base[ev, 0]).max() < 2e-5
    ctx.close()


def test_query_large_leaves(cabi, oracle, oracle64):
    """Leaf systems beyond the 8-query kernel's shared-memory budget: n ~ 1400 runs the 4-query variant,
    n ~ 2700 the one-CTA-per-pair kernel; both against the oracle."""
    rng = np.random.default_rng(9)
    P = H.P3
    ctx = cabi.Ctx(3)
    cells = np.array([[3, 1, -2], [9, 1, -2], [15, 1, -2], [21, 1, -2]], np.int32)
    centres = ((2 * cells + 1) * np.float64(np.float32(P["half"]))).astype(np.float32)
    sizes = [120, 360, 500, 740]   # n ~ 430 (8-query kernel), ~1290 (6-query), ~1790 (4-query), ~2650 (per-pair kernel)
    offs, chunks = [0], []
    for c, N in zip(centres, sizes):
        s = H.leaf_samples3(N, rng, spread=0.045)
        s[:, :3] += c - np.array([0.3125, -0.1375, 0.0625], np.float32)
        s[:, 8] = rng.uniform(0.01, 0.08, N)
        chunks.append(s)
        offs.append(offs[-1] + N)
    samples = np.concatenate(chunks)
    st = ctx.leaves_update(cells, centres, offs, samples)
    assert (st == 0).all()
    ns = [ctx.leaf_get(c, want_L=False)["n"] for c in cells]
    assert 1280 < ns[1] <= 1696 < ns[2] <= 2560 < ns[3]
    x = np.concatenate([c + rng.uniform(-0.02, 0.02, (40, 3)) for c in centres]).astype(np.float32)
    got = ctx.query(x)
    order = np.argsort(cells[:, 0])     # same y, z: DFS order = x ascending
    gps = [oracle.gp_train(3, chunks[i], P["scale"], P["noise"]) for i in order]
    gps64 = [oracle64.gp_train(3, chunks[i], P["scale"], P["noise"]) for i in order]
    want = oracle.make_map(3, centres[order], P["half"], gps, P["search"], P["var_thre"], P["noise"]).test(x)
    want64 = oracle64.make_map(3, centres[order], P["half"], gps64, P["search"], P["var_thre"], P["noise"]).test(x)
    # 740 samples inside one 5 cm ball are nearly duplicated: fp32 pins fewer rows than on real maps
    H.check_rows(got, want, want64, 3, label="large leaves", max_unpinned=0.4)
    ctx.close()


def test_snapshot_roundtrip(cabi, tmp_path):
    """f-4: the trained device map written to a flat file (the message gpis_replicate ships, with every leaf in it)
    and loaded into a fresh context answers bit-identically, including leaves registered without a GP, effective
    boxes and the root box (tie order). A file with different map parameters is refused."""
    ctx, g, cells, P = _fixture_map(cabi)
    ctx.leaves_mark([[40, 40, 40]], [[2.025, 2.025, 2.025]])           # registered, untrained
    path = str(tmp_path / "map.gpis")
    ctx.snapshot_save(path)
    assert os.path.getsize(path) > 1 << 20
    other = cabi.Ctx(3)
    other.leaves_update([[7, 7, 7]], [[0.375, 0.375, 0.375]], [0, 30], H.leaf_samples3(30, np.random.default_rng(1)))   # wiped by the load
    other.snapshot_load(path)
    sa, sb = ctx.stats(), other.stats()
    assert sa["leaves"] == sb["leaves"] and sa["leaves_trained"] == sb["leaves_trained"] and other.leaf_index([7, 7, 7]) == -1
    a, ca, ta = ctx.query(g["X"], g["init"].copy(), debug=True)
    b, cb, tb = other.query(g["X"], g["init"].copy(), debug=True)
    assert np.array_equal(a, b) and np.array_equal(ca[:, 0], cb[:, 0]) and np.array_equal(ta, tb)
    two = cabi.Ctx(2)
    with pytest.raises(RuntimeError):
        two.snapshot_load(path)
    two.close()
    other.close()
    ctx.close()


def test_replicate_two_gpus(cabi):
    """K5 on hardware: two ranks (one per GPU) under torch.distributed.run; rank 0 maps frames and calls
    gpis_replicate after each, rank 1 follows. Both answer the same queries bit-identically, including after erases
    and box changes (scripts/replicate_check.py). Skipped on a one-GPU box."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(root, "scripts", "replicate_check.py"), "4"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "REPLICATE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_obs_gp_matches_oracle(cabi, oracle):
    """K2: partition, per-tile training, batched test incl. margins / invalid tiles / untouched val."""
    g = dict(np.load(os.path.join(G, "obs2d.npz")))
    ctx = cabi.Ctx(3)
    ctx.obs_train_2d(g["vu"], g["zinv"], int(g["ni"]), int(g["nj"]))
    val, var = ctx.obs_test(g["xt"], 2, val=g["val0"])
    ev = g["var"] < 1e5
    assert np.array_equal(var > 1e5, ~ev)                         # same evaluated set
    assert np.array_equal(val[~ev], g["val0"][~ev])
    # K2 follows the oracle's operation order (obs_gp.cuh): bit-identical outputs
    assert np.array_equal(val[ev], g["val"][ev])
    assert np.array_equal(var[ev], g["var"][ev])
    s = dict(np.load(os.path.join(G, "seq2d.npz")))
    c2 = cabi.Ctx(2)
    f = (1.0 / np.sqrt(s["ranges"][0])).astype(np.float32)
    c2.obs_train_1d(s["thetas"], f)
    val, var = c2.obs_test(s["obs1_xt"], 1)
    ev = s["obs1_var"] < 1e5
    assert np.array_equal(var > 1e5, ~ev)
    assert np.array_equal(val[ev], s["obs1_val"][ev])
    assert np.array_equal(var[ev], s["obs1_var"][ev])
    ctx.close()
    c2.close()


def _rows_report(rows, ref, dim, label):
    """Error statistics of result rows against the reference's rows (same floors as helpers.check_rows)."""
    w = 1 + dim
    ev = ref[:, w] < 1.0
    assert np.array_equal(rows[:, w] < 1.0, ev), f"{label}: evaluated mask differs"
    ef, eg, evr = H._errs(rows[ev], ref[ev], dim)
    rep = {"rows": int(ev.sum())}
    for name, e, tol in (("f", ef, 1e-4), ("grad", eg, 1e-4), ("var", evr, 1e-3)):
        rep[name] = {"median": float(np.median(e)), "p99": float(np.percentile(e, 99)), "max": float(e.max()),
                     "within_tol": float((e < tol).mean())}
    print(label, rep)
    return rep


def test_gpismap2d_end_to_end(cabi):
    """The drop-in GPisMap on the first laser scans of the bundled sequence (stored in the fixture): after every
    scan the leaves AND the samples (position, normal, value, both noises) are bit-identical to the reference's —
    the observation GP follows the oracle's operation order — and the result rows agree within tolerance."""
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "seq2d.npz")))
    m = hostapi.GPisMap()
    assert m.test(g["X"]) is None
    for i in range(g["ranges"].shape[0]):
        m.update(g["thetas"], g["ranges"][i], g["pose6"][i])
        c, n = m.leaves()
        assert np.array_equal(c, g[f"leaves{i}"])                 # leaf assignment bit-exact
        assert np.array_equal(n, g[f"leafcount{i}"])
        assert np.array_equal(m.all_samples(), g[f"samples{i}"]), f"samples differ after scan {i}"
    rep = _rows_report(m.test(g["X"]), g["rows"], 2, "seq2d")
    assert rep["f"]["within_tol"] >= 0.97 and rep["f"]["p99"] < 1e-3, rep
    assert rep["var"]["within_tol"] >= 0.97, rep
    m.close()


def test_gpismap3_synthetic_frames(cabi, oracle):
    """GPisMap3 on two synthetic depth frames: every class method, timing counters, and the map it
    trains answers queries like the oracle trained on the same samples."""
    from gpismap_b200 import hostapi, synth
    m = hostapi.GPisMap3()
    for k in range(2):
        dz, pose = synth.frame(k, 40)
        m.update(dz, pose)
        ph, cnt, ms = m.timing()
        # training overlaps the next frame (gpis_set_train_mode): the reported kernel time is the last COMPLETED batch
        assert cnt[0] > 70000 and cnt[2] > 0 and (ms > 0 or k == 0)
    pts = m.getAllPoints()
    assert pts.shape[0] > 80000
    assert (pts > synth.ROOM_LO - 0.05).all() and (pts < synth.ROOM_HI + 0.05).all()
    S = m.all_samples()
    nrm = np.linalg.norm(S[:, 3:6], axis=1)
    assert np.median(np.abs(nrm - 1)) < 1e-3                      # unit normals from evalPoints
    rng = np.random.default_rng(3)
    X = (S[::40, :3] - S[::40, 3:6] * rng.uniform(0.002, 0.02, (len(S[::40]), 1))).astype(np.float32)   # 2..20 mm inside the room
    rows = m.test(X)
    ev = rows[:, 4] < 0.3
    assert ev.sum() > 0.5 * len(X)
    # field sanity: near the surface the predicted gradient is (anti)parallel to the stored surface normal
    assert np.isfinite(rows).all()
    g = rows[ev, 1:4]
    nrm_s = S[::40, 3:6][ev]
    cosang = np.abs((g * nrm_s).sum(1)) / np.maximum(np.linalg.norm(g, axis=1) * np.linalg.norm(nrm_s, axis=1), 1e-9)
    assert np.median(cosang) > 0.95
    m.reset()
    assert m.getAllPoints().shape[0] == 0 and m.test(X) is None
    m.close()


def test_train_modes_give_identical_maps(cabi, monkeypatch):
    """Overlapped leaf training (gpis_set_train_mode 1, 2 and 3, include/gpis_b200.h) never changes a result: the samples
    after four frames and the answers to the same queries are bit-identical to the synchronous mode, queries issued
    right after update() see the batch that is still in flight, and the training time is reported once it completed."""
    from gpismap_b200 import hostapi, synth
    out = {}
    for mode in (0, 1, 2, 3):
        monkeypatch.setenv("GPIS_TRAIN_MODE", str(mode))
        m = hostapi.GPisMap3()
        mids = []
        for k in range(4):
            dz, pose = synth.frame(k, 40)
            m.update(dz, pose)
            if k == 1:
                S1 = m.all_samples()
                Xm = (S1[::97, :3] - S1[::97, 3:6] * np.float32(0.01)).astype(np.float32)
                mids = m.test(Xm)                      # waits for (mode 2: also launches) the batch of frame 1
        S = m.all_samples()
        X = (S[::53, :3] - S[::53, 3:6] * np.float32(0.008)).astype(np.float32)
        rows = m.test(X)
        ph, cnt, ms = m.timing()
        out[mode] = (S, mids, rows, cnt[2])
        m.close()
    for mode in (1, 2, 3):
        assert np.array_equal(out[mode][0], out[0][0]), mode
        assert np.array_equal(out[mode][1], out[0][1], equal_nan=True), mode
        assert np.array_equal(out[mode][2], out[0][2], equal_nan=True), mode
        assert out[mode][3] == out[0][3] > 0


def test_bench_scale_map_matches_reference(cabi):
    """End to end at the bench's scale (BASELINE configs[1] + [2], leaf systems of n ~ 1,100-1,600 unknowns):
    the drop-in GPisMap3 maps the 40 synthetic frames through the GPU; its complete sample array (508,149 samples)
    must hash to what the unmodified reference produced when it mapped the same frames itself
    (tests/golden/room40.npz). Then the frozen region of the reference's own octree is loaded into oracle/_ref, its
    updateGPs + test run there, and the CUDA path (all leaves retrained on the final samples, like the reference
    side) must agree on the grid points of that region: identical evaluated mask, and with the fp64 evaluation of the
    same formulas as arbiter (helpers.check_rows semantics) every row the fp32 reference pins is within the
    north_star tolerance (1e-4 on f and grad f, 1e-3 on the variances), every other row no further from fp64 than 4x
    the reference's own distance."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import refpy
    if not refpy.available():
        pytest.skip("oracle/_ref (the compiled reference) is not present")
    import bench
    from gpismap_b200 import hostapi, synth
    m = hostapi.GPisMap3()
    g = np.load(os.path.join(G, "room40.npz"))
    for k in range(40):
        dz, pose = synth.frame(k, 40)
        m.update(dz, pose)
        assert m.getAllPoints().shape[0] == g["nsamples"][k], f"sample count differs after frame {k}"

    class Args:
        cpu_baseline_seconds = 2.0
        frames = 40
        noise_mm = 1.0
        grid = 256
        parity_half = 0.2
        parity_rows = 6000
    out = bench.cpu_baseline(Args, m, None)
    m.close()
    p = out["parity_vs_reference"]
    print("bench-scale parity:", p)
    assert "error" not in p, p
    assert p["gpu_map_samples_sha256_equals_reference_map"], p
    assert p["evaluated_mask_identical"] and p["oracle_fp32_equals_reference_rows"], p
    a = p["fp64_arbitration"]
    assert a["rows"] > 1000, p
    for name in ("f", "grad", "var"):
        assert a[name]["pinned_rows"] > 0 and a[name]["pinned_within_tol"] == 1.0, (name, a[name])
        assert a[name]["unpinned_no_further_from_fp64_than_4x_reference"] == a[name]["unpinned_rows"], (name, a[name])
    r = p["all_rows"]
    assert r["f_rel"]["within_1e-4"] >= 0.999 and r["var_rel"]["within_1e-3"] >= 0.999, r


def test_query_on_sample_position_nan_pattern(cabi, oracle):
    """r = 0 makes kf2 divide by zero in the reference (covFnc.cpp:29-33): a query placed exactly on a sample with
    a usable normal gets NaN gradients and gradient variances there, finite f and var_f. The CUDA path must
    reproduce the same pattern, and agree on the finite entries."""
    rng = np.random.default_rng(2)
    P = H.P3
    ctx = cabi.Ctx(3)
    cells = np.array([[6, -3, 1]], np.int32)
    centres = ((2 * cells + 1) * np.float64(np.float32(P["half"]))).astype(np.float32)
    s = H.leaf_samples3(60, rng, spread=0.03)
    s[:, :3] += centres[0] - np.array([0.3125, -0.1375, 0.0625], np.float32)
    assert (ctx.leaves_update(cells, centres, [0, len(s)], s) == 0).all()
    x = np.concatenate([s[:12, :3], s[12:16, :3] + np.float32(1e-3)]).astype(np.float32)
    got = ctx.query(x)
    gp = oracle.gp_train(3, s, P["scale"], P["noise"])
    want = oracle.make_map(3, centres, P["half"], [gp], P["search"], P["var_thre"], P["noise"]).test(x)
    assert np.isnan(want).any() and not np.isnan(want[:, 0]).any()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    fin = ~np.isnan(want)
    assert np.allclose(got[fin], want[fin], rtol=2e-3, atol=1e-5)
    ctx.close()


def test_gpismap2d_whole_demo_sequence(cabi):
    """BASELINE configs[0]: all 28 scans of the reference's 2-D demo (matlab/demo_gpisMap.m) through the drop-in
    GPisMap: the leaf count after every scan, the final leaf set and the final samples are bit-identical to the
    reference's; the evaluated mask on the demo grid is identical and the rows agree within tolerance."""
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "seq2d_demo.npz")))
    m = hostapi.GPisMap()
    for i in range(g["ranges"].shape[0]):
        m.update(g["thetas"], g["ranges"][i], g["pose6"][i])
        assert m.leaves()[0].shape[0] == g["nleaves"][i], i
    assert np.array_equal(m.leaves()[0], g["leaves"])
    assert np.array_equal(m.all_samples(), g["samples"])
    rep = _rows_report(m.test(g["X"]), g["rows"], 2, "seq2d_demo")
    assert rep["rows"] > 1000
    assert rep["f"]["within_tol"] >= 0.97 and rep["f"]["p99"] < 1e-3, rep
    m.close()


def test_query2d_matches_reference_fixture(cabi, oracle, oracle64r):
    """K3 + K4 in 2-D against rows produced by the unmodified reference (tests/golden/map2d.npz,
    GPisMap::test_kernel, GPisMap.cpp:665-763): candidate counts and picks bit-exact (lattice-plane queries, exact
    ties, > 16 candidates), rows within the stated tolerances with fp64 arbitration."""
    g = dict(np.load(os.path.join(G, "map2d.npz")))
    P = H.P2
    ctx = cabi.Ctx(2)
    pitch = 2.0 * np.float64(np.float32(P["half"]))
    root_min = np.round((g["root_c"].astype(np.float64) - float(g["root_half"])) / pitch).astype(np.int32)
    levels = int(round(np.log2(float(g["root_half"]) / np.float64(np.float32(P["half"])))))
    ctx.rebase(root_min, levels)
    cells = cabi.cells_of(g["centres"], P["half"])
    st = ctx.leaves_update(cells, g["centres"], g["offsets"], g["samples"])
    assert (st == 0).all()
    ctx.leaves_set_boxes(cells, g["boxes"])
    got, chosen, tie = ctx.query(g["X"], g["init"].copy(), debug=True)
    assert np.array_equal(chosen[:, 0], g["ncand"])
    offs = g["offsets"]
    gps = [oracle.gp_train(2, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    gps64 = [oracle64r.gp_train(2, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    m = oracle.make_map(2, g["centres"], P["half"], gps, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    m64 = oracle64r.make_map(2, g["centres"], P["half"], gps64, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    want, ochosen, otie = m.test(g["X"], g["init"].copy(), want_choice=True)
    want64 = m64.test(g["X"], g["init"].astype(np.float64))
    assert np.array_equal(want, g["rows"])                      # the oracle itself reproduces the fixture
    slot = np.array([ctx.leaf_index(c) for c in cells])
    inv = -np.ones(slot.max() + 2, np.int64)
    inv[slot] = np.arange(len(slot))
    ch = chosen.copy()
    for k in (1, 2, 3):
        ch[:, k] = np.where(chosen[:, k] >= 0, inv[np.maximum(chosen[:, k], 0)], -1)
    assert np.array_equal(ch, ochosen), "neighbour picks differ from the reference's std::sort order"
    assert np.array_equal(tie, otie)
    ev = g["ncand"] > 0
    evi = np.flatnonzero(ev)
    rep = H.check_rows(got[ev], g["rows"][ev], want64[ev], 2, label="map2d",
                       explain=lambda i: H.selection_ambiguity(gps64, ochosen[evi[i]], g["X"][evi[i]], P["var_thre"], 2))
    print("map2d", rep)
    keep = [0, 1, 2, 4, 5]
    assert np.array_equal(got[~ev][:, keep], g["init"][~ev][:, keep])
    ctx.close()


def test_gpismap3_bundled_sequence(cabi):
    """The reference's own 3-D demo run (matlab/demo_gpisMap3.m: 40 masked BigBIRD depth frames, intrinsics switched
    per frame at a fixed resolution, which exercises the stale ObsGP2D partition of SURVEY 9-12) through the
    drop-in GPisMap3: leaf and sample counts after every frame, the samples after frames 5, 20 and 40 and the
    final leaf set are bit-identical to the reference's; rows on the demo grid within tolerance."""
    from gpismap_b200 import hostapi
    BIGBIRD_CAMS = H.BIGBIRD_CAMS
    g = dict(np.load(os.path.join(G, "seq3d.npz")))
    m = None
    for k in range(len(g["cam"])):
        cam = int(g["cam"][k])
        c = tuple(np.float32(BIGBIRD_CAMS[n][cam - 1]) for n in ("fx", "fy", "cx", "cy")) + (640, 480)
        if m is None:
            m = hostapi.GPisMap3(cam=c)
        else:
            m.resetCam(*c)
        dz = np.zeros(640 * 480, np.float32)
        a, b = g["depth_off"][k], g["depth_off"][k + 1]
        dz[g["depth_idx"][a:b]] = g["depth_val"][a:b]
        m.update(dz, g["pose12"][k])
        S = m.all_samples()
        assert len(S) == g["nsamples"][k] and m.leaves()[0].shape[0] == g["nleaves"][k], (k, len(S), g["nsamples"][k])
        if f"samples{k + 1}" in g:
            assert np.array_equal(S, g[f"samples{k + 1}"]), f"samples differ after frame {k + 1}"
    assert np.array_equal(m.leaves()[0], g["leaves"])
    rep = _rows_report(m.test(g["X"]), g["rows"], 3, "seq3d")
    assert rep["rows"] > 4000
    assert rep["f"]["within_tol"] >= 0.97 and rep["f"]["p99"] < 1e-3, rep
    m.close()
