"""Host-side logic of the drop-in classes (tree, evalPoints / reEvalPoints, dirty-leaf CSR, table
sync, effective boxes, DFS tie order) on CPU: the classes are linked against tests/mock_cabi — the
same C ABI backed by the oracle — and must reproduce the reference's golden outputs bit for bit."""
import os

import numpy as np
import pytest

import helpers as H

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mocklib():
    import mockbuild
    return mockbuild.build()


def test_2d_sequence_bit_exact(mocklib):
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "seq2d.npz")))
    m = hostapi.GPisMap(libpath=mocklib)
    assert m.test(g["X"]) is None                       # test before update: false, not a null dereference
    for i in range(g["ranges"].shape[0]):
        m.update(g["thetas"], g["ranges"][i], g["pose6"][i])
        assert np.array_equal(m.all_samples(), g[f"samples{i}"])        # leaf assignment, normals, noise
        c, n = m.leaves()
        assert np.array_equal(c, g[f"leaves{i}"]) and np.array_equal(n, g[f"leafcount{i}"])
    assert np.array_equal(m.test(g["X"]), g["rows"])     # includes lattice-plane and exact-tie queries
    m.reset()
    assert m.getAllPoints().shape[0] == 0 and m.test(g["X"]) is None


def test_bad_arguments_return_silently(mocklib):
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "seq2d.npz")))
    m = hostapi.GPisMap(libpath=mocklib)
    m.update(g["thetas"], np.zeros_like(g["ranges"][0]), g["pose6"][0])   # no valid range: ignored
    assert m.getAllPoints().shape[0] == 0


def test_3d_bulk_load_matches_reference_fixture(mocklib):
    """insertSamples + trainActive against the map3d fixture: same leaves (DFS order), same rows."""
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "map3d.npz")))
    m = hostapi.GPisMap3(libpath=mocklib)
    assert m.insert_samples(g["samples_in"]) == len(g["samples_in"])
    m.train_active()
    c, n = m.leaves()
    assert np.array_equal(c, g["centres"])
    assert np.array_equal(m.test(g["X"], g["init"].copy()), g["rows"])


def test_tree_against_reference(ref, mocklib):
    """Random insert streams incl. duplicates, lattice-plane points and root growth: same surviving samples."""
    from gpismap_b200 import hostapi
    rng = np.random.default_rng(21)
    pts = rng.uniform(-0.45, 0.45, (4000, 3))
    pts[::50] = np.round(pts[::50] / 0.003125) * 0.003125          # on lattice planes: dropped
    pts[1::60] = pts[0:1] + rng.normal(0, 1e-3, (len(pts[1::60]), 3))   # closer than the minimum spacing
    pts[2::97] += 0.9                                               # outside the initial root: root growth
    s = np.zeros((len(pts), 9), np.float32)
    s[:, :3] = pts
    s[:, 5] = 1
    s[:, 6] = -0.2
    s[:, 7] = 0.003
    s[:, 8] = 0.02
    R = ref.RefMap3()
    M = hostapi.GPisMap3(libpath=mocklib)
    assert R.insert_samples(s) == M.insert_samples(s)
    assert np.array_equal(R.all_samples(), M.all_samples())
    ca, na, _ = R.clusters()
    cb, nb = M.leaves()
    assert np.array_equal(ca, cb) and np.array_equal(na, nb)


def test_3d_bundled_sequence_bit_exact(mocklib):
    """First frames of the reference's 3-D demo run (tests/golden/seq3d.npz) through the drop-in GPisMap3 on the
    oracle-backed mock ABI: sample and leaf counts after every frame and the samples after frame 5 are the
    reference's, bit for bit (preprocData, regressObs, reEvalPoints, evalPoints, tree logic; intrinsics change per
    frame at fixed resolution: the stale ObsGP2D partition of SURVEY 9-12 is part of the contract)."""
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "seq3d.npz")))
    m = None
    for k in range(6):
        cam = int(g["cam"][k])
        c = tuple(np.float32(H.BIGBIRD_CAMS[n][cam - 1]) for n in ("fx", "fy", "cx", "cy")) + (640, 480)
        if m is None:
            m = hostapi.GPisMap3(cam=c, libpath=mocklib)
        else:
            m.resetCam(*c)
        dz = np.zeros(640 * 480, np.float32)
        a, b = g["depth_off"][k], g["depth_off"][k + 1]
        dz[g["depth_idx"][a:b]] = g["depth_val"][a:b]
        m.update(dz, g["pose12"][k])
        S = m.all_samples()
        assert len(S) == g["nsamples"][k] and m.leaves()[0].shape[0] == g["nleaves"][k], k
        if k + 1 == 5:
            assert np.array_equal(S, g["samples5"])


def test_3d_larger_training_balls_match_reference_variant(mocklib):
    """BASELINE configs[4] knob: GPisMap3Tuning.rtimes = 2.5 (the reference needs a rebuild with GPISMAP3_RTIMES 2.5,
    oracle/params_variants/rtimes25). Fixture from that variant build: same leaves, same training sets (sizes), same rows."""
    from gpismap_b200 import hostapi
    g = dict(np.load(os.path.join(G, "map3d_rt25.npz")))
    m = hostapi.GPisMap3(libpath=mocklib, rtimes=2.5)
    assert m.insert_samples(g["samples_in"]) == len(g["samples_in"])
    m.train_active()
    c, n = m.leaves()
    assert np.array_equal(c, g["centres"])
    assert np.array_equal(m.test(g["X"], g["init"].copy()), g["rows"])
    d = hostapi.GPisMap3(libpath=mocklib)            # default radius: different training sets, different rows
    d.insert_samples(g["samples_in"])
    d.train_active()
    assert not np.array_equal(d.test(g["X"], g["init"].copy()), g["rows"])
