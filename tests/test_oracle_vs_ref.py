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
