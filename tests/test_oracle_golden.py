"""The plain-C restatement oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). CPU only; runs wherever the repo is checked out."""
import os

import numpy as np

import helpers as H

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


def _leaf(oracle, dim, name):
    g = load(name)
    P = H.P3 if dim == 3 else H.P2
    gp = oracle.gp_train(dim, g["samples"], P["scale"], P["noise"])
    alpha, L, gf = gp.factors()
    assert np.array_equal(gf, g["gradflag"])                     # gradflag rule, OnGPIS.cpp:63-66 / 122-125
    sx = np.where(gf > 0, g["samples"][:, 2 * dim + 1], 2.0).astype(np.float32)
    K = oracle.matern_train(dim, g["samples"][:, :dim], gf, P["scale"], sx, g["samples"][:, 2 * dim + 2])
    assert np.array_equal(K, g["K"])                             # covariance entries bit-exact
    assert np.array_equal(K, K.T)
    # same fp32 operation order as the shim-backed reference build: identical factors and predictions
    assert np.array_equal(alpha, g["alpha"])
    assert np.array_equal(L, g["L"])
    assert np.array_equal(gp.test(g["x"]), g["rows"])
    assert gp.chol_fail == 0


def test_leaf3d_matches_reference(oracle):
    _leaf(oracle, 3, "leaf3d.npz")


def test_leaf2d_matches_reference(oracle):
    _leaf(oracle, 2, "leaf2d.npz")


def test_leaf_vs_fp64_shadow(oracle, oracle64):
    g = load("leaf3d.npz")
    a = oracle.gp_train(3, g["samples"], 0.04, 5e-3).test(g["x"])
    b = oracle64.gp_train(3, g["samples"], 0.04, 5e-3).test(g["x"].astype(np.float64))
    assert np.abs(a[:, 0] - b[:, 0]).max() < 1e-5
    assert (np.abs(a[:, 4:] - b[:, 4:]) / np.maximum(np.abs(b[:, 4:]), 1e-3)).max() < 1e-3


def test_obs2d_matches_reference(oracle):
    g = load("obs2d.npz")
    o = oracle.obs2d(g["vu"], g["zinv"], int(g["ni"]), int(g["nj"]))
    bi, bj = o.bounds()
    assert np.array_equal(bi, g["bi"]) and np.array_equal(bj, g["bj"])    # partition, ObsGP.cpp:204-265
    assert np.array_equal(o.tile_counts(), g["tiles"])
    val, var = o.test(g["xt"], val=g["val0"])
    assert np.array_equal(var, g["var"])
    assert np.array_equal(val, g["val"])
    skipped = g["var"] > 1e5
    assert skipped.any() and np.array_equal(val[skipped], g["val0"][skipped])   # val untouched, var = 1e6


def test_obs1d_matches_reference(oracle):
    g = load("seq2d.npz")
    f = (1.0 / np.sqrt(g["ranges"][0])).astype(np.float32)
    o = oracle.obs1d(g["thetas"], f)
    r, _ = o.bounds()
    assert np.array_equal(r, g["obs1_ranges"])
    assert len(o.tile_counts()) == 14 and o.tile_counts().tolist() == [26] * 12 + [22, 15]
    val, var = o.test(g["obs1_xt"])
    assert np.array_equal(val, g["obs1_val"]) and np.array_equal(var, g["obs1_var"])


def test_map3d_query_matches_reference(oracle):
    """Candidate rule, std::sort replay on exact ties (> 16 candidates) and fusion: bit-identical rows."""
    g = load("map3d.npz")
    P = H.P3
    offs = g["offsets"]
    gps = [oracle.gp_train(3, g["samples"][offs[i]:offs[i + 1]], P["scale"], P["noise"]) for i in range(len(offs) - 1)]
    m = oracle.make_map(3, g["centres"], P["half"], gps, P["search"], P["var_thre"], P["noise"], boxes=g["boxes"])
    rows, chosen, tie = m.test(g["X"], g["init"].copy(), want_choice=True)
    assert np.array_equal(chosen[:, 0], g["ncand"])            # neighbour counts bit-exact
    assert (tie > 0).sum() > 50 and (g["ncand"] > 16).sum() > 50   # the fixture does exercise the tie path
    assert np.array_equal(rows, g["rows"])
    untouched = g["ncand"] == 0
    assert untouched.any()
    keep = [0, 1, 2, 3, 5, 6, 7]
    assert np.array_equal(rows[untouched][:, keep], g["init"][untouched][:, keep])   # read-modify-write semantics
    assert np.all(rows[untouched][:, 4] == np.float32(1.0 + np.float32(5e-3)))


def test_closed_forms(oracle):
    """kf(0)=1 on the diagonal before noise, single no-gradient sample GP."""
    s = np.zeros((1, 9), np.float32)
    s[0, :3] = (0.1, 0.2, 0.3)
    s[0, 6] = -0.2
    s[0, 7] = 0.004       # replaced by 2.0 because the normal is null
    gp = oracle.gp_train(3, s, 0.04, 5e-3)
    alpha, L, gf = gp.factors()
    assert gp.n == 1 and gf[0] == 0
    assert np.isclose(L[0, 0] ** 2, 3.0) and np.isclose(alpha[0], -0.2 / 3.0)
    x = np.array([[0.1, 0.2, 0.34]], np.float32)
    r = np.float32(np.linalg.norm(x[0] - s[0, :3]))
    a = np.sqrt(3.0) / 0.04
    kf = (1 + a * r) * np.exp(-a * r)
    row = gp.test(x)[0]
    assert np.isclose(row[0], kf * (-0.2 / 3.0), rtol=1e-5)
    assert np.isclose(row[4], 1.001 - kf * kf / 3.0, rtol=1e-5)
