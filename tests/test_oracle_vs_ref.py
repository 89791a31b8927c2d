"""The C restatement oracle against the unmodified reference built into oracle/_ref (skipped where
that build is absent). Larger, randomised cases than the committed golden vectors."""
import numpy as np
import pytest

import helpers as H


@pytest.mark.parametrize("dim,N", [(3, 1), (3, 7), (3, 60), (3, 150), (2, 1), (2, 9), (2, 75)])
def test_leaf_gp(ref, oracle, dim, N):
    rng = np.random.default_rng(100 + N)
    P = H.P3 if dim == 3 else H.P2
    s = H.leaf_samples3(N, rng) if dim == 3 else H.leaf_samples2(N, rng)
    r = ref.RefGP(dim, s, P["scale"], P["noise"])
    o = oracle.gp_train(dim, s, P["scale"], P["noise"])
    ar, Lr, gfr = r.factors()
    ao, Lo, gfo = o.factors()
    assert r.n == o.n and np.array_equal(gfr, gfo)
    assert np.array_equal(ar, ao) and np.array_equal(Lr, Lo)
    x = (s[:, :dim] + rng.normal(0, P["half"] * 0.4, (N, dim))).astype(np.float32)
    assert np.array_equal(r.test(x), o.test(x))


def test_obs2d_full_frame(ref, oracle):
    rng = np.random.default_rng(3)
    ni, nj = 60, 80
    v = ((np.arange(ni) * 2 - 50) / 568.0).astype(np.float32)
    u = ((np.arange(nj) * 2 - 70) / 568.0).astype(np.float32)
    vu = np.zeros((nj, ni, 2), np.float32)
    vu[:, :, 0] = v[None, :]
    vu[:, :, 1] = u[:, None]
    zinv = (1.0 / (1.0 + 0.5 * np.sin(9 * vu[:, :, 0]) ** 2 + vu[:, :, 1])).astype(np.float32)
    zinv[rng.uniform(size=zinv.shape) < 0.2] = -1
    r = ref.RefObs2D()
    r.train(vu, zinv, ni, nj)
    o = oracle.obs2d(vu, zinv, ni, nj)
    xt = np.stack([rng.uniform(v[0], v[-1], 500), rng.uniform(u[0], u[-1], 500)], 1).astype(np.float32)
    rv, rr = r.test(xt)
    ov, orr = o.test(xt)
    assert np.array_equal(rv, ov) and np.array_equal(rr, orr)
    assert np.array_equal(r.tiles(), o.tile_counts())


@pytest.mark.parametrize("dim", [3, 2])
def test_map_query(ref, oracle, dim):
    rng = np.random.default_rng(7)
    P = H.P3 if dim == 3 else H.P2
    if dim == 3:
        M = ref.RefMap3()
        s = H.sphere_samples(0.15, 0.0125, (0.0517, 0.0231, 0.0113), rng)
    else:
        M = ref.RefMap2()
        s = H.circle_samples(6.3, 0.45, (1.7, -2.3), rng)
    assert M.insert_samples(s) == len(s)
    M.update_gps()
    centres, offs, samples, trained = H.ref_map_to_csr(M, P)
    assert (trained == np.diff(offs)).all()
    gps = [oracle.gp_train(dim, samples[offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(centres))]
    m = oracle.make_map(dim, centres, P["half"], gps, P["search"], P["var_thre"], P["noise"], boxes=M.cluster_boxes())
    n = 1500
    d = rng.normal(size=(n, dim))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if dim == 3:
        x = d * (0.15 + rng.uniform(-0.1, 0.1, n))[:, None] + np.array([0.0517, 0.0231, 0.0113])
    else:
        x = d * (6.3 + rng.uniform(-6, 6, n))[:, None] + np.array([1.7, -2.3])
    # add lattice-aligned points (exact ties, touching boxes)
    lat = np.round(x[:300] / P["half"]) * P["half"]
    x = np.concatenate([x, lat]).astype(np.float32)
    init = rng.uniform(size=(len(x), 2 * (1 + dim))).astype(np.float32)
    want = M.test(x, init.copy())
    got, chosen, tie = m.test(x, init.copy(), want_choice=True)
    nc = np.array([M.candidates(q, P["search"])[0].shape[0] for q in x])
    assert np.array_equal(chosen[:, 0], nc)
    assert np.array_equal(got, want)


def test_std_sort_replay_matches_libstdcxx_including_heapsort_fallback(ref, oracle):
    """The candidate order is whatever libstdc++'s std::sort leaves (GPisMap3.cpp:826-829): not stable, so exact ties
    are ordered by its partitioning. The oracle (and the CUDA kernel, same code) replays it move by move; here the
    replay is pinned against the real std::sort on tie-heavy random keys, on sorted / reversed / organ-pipe inputs and
    on McIlroy-adversary inputs that push the introsort past its depth limit into the heapsort fallback."""
    rng = np.random.default_rng(7)
    cases = []
    for n in (1, 2, 3, 16, 17, 18, 31, 32, 33, 64, 100, 125, 200, 256):
        cases.append(rng.integers(0, 6, n).astype(np.float32))            # many exact ties
        cases.append(rng.integers(0, max(2, n // 2), n).astype(np.float32))
        cases.append(rng.uniform(size=n).astype(np.float32))
        cases.append(np.arange(n, dtype=np.float32))
        cases.append(np.arange(n, dtype=np.float32)[::-1].copy())
        cases.append(np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.float32))
        if n > 16:
            cases.append(ref.std_sort_killer(n))
            cases.append(ref.std_sort_killer(n, ties=2))               # the adversary's shape with pairs of exact ties
            cases.append(ref.std_sort_killer(n, ties=3))
    h0 = oracle.L.gpo_sort_replay_heapsorts()
    for k in cases:
        a = ref.std_sort_indices(k)
        b = oracle.sort_replay(k)
        assert np.array_equal(a, b), (len(k), k[:20])
        assert np.all(np.diff(k[a]) >= 0)
    # the adversarial inputs really tripped the depth limit, with ties present (so the fallback's move order matters)
    assert oracle.L.gpo_sort_replay_heapsorts() - h0 >= 10
