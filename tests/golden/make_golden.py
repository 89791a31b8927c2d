"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgpisref.so, built from
/root/reference by oracle/Makefile) — run in the build container where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these fixtures are outputs of the
reference's own code on small seeded inputs, and pin the C restatement oracle on machines where
oracle/_ref cannot be rebuilt. Bundled-data inputs (three laser scans of data/2D/gazebo1.mat) are
stored next to the outputs so nothing reads /root/reference at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle import refpy  # noqa: E402


def leaf(dim, N, seed):
    rng = np.random.default_rng(seed)
    P = H.P3 if dim == 3 else H.P2
    s = H.leaf_samples3(N, rng) if dim == 3 else H.leaf_samples2(N, rng)
    g = refpy.RefGP(dim, s, P["scale"], P["noise"])
    alpha, L, gf = g.factors()
    sx = np.where(gf > 0, s[:, 2 * dim + 1], 2.0).astype(np.float32)
    K = refpy.matern_train(dim, s[:, :dim], gf, P["scale"], sx, s[:, 2 * dim + 2])
    if dim == 3:
        x = (s[::2, :3] + rng.normal(0, 0.01, s[::2, :3].shape)).astype(np.float32)
    else:
        x = (s[::2, :2] + rng.normal(0, 0.3, s[::2, :2].shape)).astype(np.float32)
    Ks = refpy.matern_test(dim, s[:, :dim], gf, x[:1], P["scale"])
    rows = g.test(x)
    return dict(samples=s, gradflag=gf, K=K, alpha=alpha, L=L, x=x, rows=rows, Ks0=Ks)


def obs2d(seed):
    rng = np.random.default_rng(seed)
    ni, nj = 23, 28
    v = ((np.arange(ni) * 2 - 20) / 568.0).astype(np.float32)
    u = ((np.arange(nj) * 2 - 31) / 568.0).astype(np.float32)
    vu = np.zeros((nj, ni, 2), np.float32)
    vu[:, :, 0] = v[None, :]
    vu[:, :, 1] = u[:, None]
    z = 1.2 + 3.0 * vu[:, :, 0] - 2.0 * vu[:, :, 1] + 0.05 * np.sin(90 * vu[:, :, 1])
    zinv = (1.0 / z).astype(np.float32)
    zinv[rng.uniform(size=zinv.shape) < 0.15] = -1.0
    zinv[:9, :9] = -1.0    # one fully invalid tile
    o = refpy.RefObs2D()
    o.train(vu, zinv, ni, nj)
    bi, bj = o.partition()
    xt = np.stack([rng.uniform(v[0] - 0.01, v[-1] + 0.01, 96), rng.uniform(u[0] - 0.01, u[-1] + 0.01, 96)], 1).astype(np.float32)
    val0 = rng.uniform(size=96).astype(np.float32)
    val, var = o.test(xt, val=val0)
    return dict(vu=vu.reshape(-1), zinv=zinv.reshape(-1), ni=ni, nj=nj, bi=bi, bj=bj, tiles=o.tiles(), xt=xt, val0=val0, val=val, var=var)


def seq2d():
    import scipy.io
    m = scipy.io.loadmat("/root/reference/data/2D/gazebo1.mat")
    thetas = m["thetas"].ravel().astype(np.float32)
    frames = [100, 200, 300, 400]
    ranges = m["ranges"][frames].astype(np.float32)
    poses = m["poses"][frames]
    pose6 = np.stack([[x, y, np.cos(p), np.sin(p), -np.sin(p), np.cos(p)] for x, y, p in poses]).astype(np.float32)
    M = refpy.RefMap2()
    xs = np.arange(-4.8, 12.1, 0.4)
    ys = np.arange(-12.0, 4.1, 0.4)    # multiples of 0.4: hits lattice planes and exact distance ties
    xg, yg = np.meshgrid(xs, ys)
    X = np.stack([xg.ravel(), yg.ravel()], 1).astype(np.float32)
    out = dict(thetas=thetas, ranges=ranges, pose6=pose6, X=X)
    for i in range(len(frames)):
        M.update(thetas, ranges[i], pose6[i])
        out[f"samples{i}"] = M.all_samples()
        c, n, tr = M.clusters()
        out[f"leaves{i}"] = c
        out[f"leafcount{i}"] = n
    out["rows"] = M.test(X)
    out["boxes"] = M.cluster_boxes()
    # 1-D observation GP on the first scan
    o = refpy.RefObs1D()
    f = (1.0 / np.sqrt(ranges[0])).astype(np.float32)
    o.train(thetas, f)
    rng = np.random.default_rng(5)
    xt = rng.uniform(thetas[0] - 0.05, thetas[-1] + 0.05, 128).astype(np.float32)
    val, var = o.test(xt)
    out.update(obs1_ranges=o.ranges(), obs1_xt=xt, obs1_val=val, obs1_var=var)
    return out


def seq2d_demo():
    """The whole 2-D demo run (matlab/demo_gpisMap.m:26-40: frames 101:100:2801 of data/2D/gazebo1.mat, test grid
    0.1 m over [-5,20]x[-15,5]): inputs for BASELINE configs[0], and the reference's leaves after every scan plus
    its result rows on every 7th grid point after the last one."""
    import scipy.io
    m = scipy.io.loadmat("/root/reference/data/2D/gazebo1.mat")
    thetas = m["thetas"].ravel().astype(np.float32)
    frames = list(range(100, m["poses"].shape[0], 100))       # 0-based 100, 200, ..., 2800
    ranges = m["ranges"][frames].astype(np.float32)
    poses = m["poses"][frames]
    pose6 = np.stack([[x, y, np.cos(p), np.sin(p), -np.sin(p), np.cos(p)] for x, y, p in poses]).astype(np.float32)
    xs = np.arange(-5 + 0.1, 20 - 0.1 + 1e-9, 0.1)
    ys = np.arange(-15 + 0.1, 5 - 0.1 + 1e-9, 0.1)
    xg, yg = np.meshgrid(xs, ys)
    X = np.stack([xg.T.ravel(), yg.T.ravel()], 1).astype(np.float32)[::7]
    M = refpy.RefMap2()
    nleaves = []
    for i in range(len(frames)):
        M.update(thetas, ranges[i], pose6[i])
        c, n, tr = M.clusters()
        nleaves.append(len(c))
    c, n, tr = M.clusters()
    return dict(thetas=thetas, ranges=ranges, pose6=pose6, X=X, rows=M.test(X), leaves=c, nleaves=np.array(nleaves, np.int32),
                nsamples=np.int32(len(M.all_samples())), samples=M.all_samples())


BIGBIRD_CAMS = H.BIGBIRD_CAMS


def seq3d():
    """The bundled 3-D demo run (matlab/demo_gpisMap3.m:26-54: 40 masked BigBIRD depth frames, camera intrinsics
    switched per frame with setCamera at a fixed 640x480 resolution, which exercises the stale ObsGP2D partition of
    SURVEY.md 9-12). Inputs are stored sparsely (valid pixels only, column-major index + metres); outputs are the
    reference's samples after frames 5, 20 and 40, its leaf count and sample count after every frame, and its result
    rows on every 3rd point of the demo's 21x25x29 test grid."""
    import cv2
    base = "/root/reference/data/3D/bigbird_detergent"
    poses = np.loadtxt(base + "/pose/poses.txt").astype(np.float32)
    frame_nums = list(range(93, 360, 3)) + list(range(3, 91, 3))
    cam_ids = [1, 2, 3, 4, 3, 2] * 30
    xg, yg, zg = np.meshgrid(np.arange(-0.07, 0.1301, 0.01), np.arange(-0.1, 0.1401, 0.01), np.arange(0, 0.2801, 0.01))
    X = np.stack([xg.ravel(order="F"), yg.ravel(order="F"), zg.ravel(order="F")], 1).astype(np.float32)[::3]
    out = dict(X=X)
    M = None
    idx_all, val_all, off, cams, pose12, nleaves, nsamples = [], [], [0], [], [], [], []
    count = 0
    for k in range(0, len(frame_nums), 3):
        frm, cam, row = frame_nums[k], cam_ids[count], poses[count]
        count += 1
        D = cv2.imread(f"{base}/masked_depth/frame{frm}_cam{cam}.png", cv2.IMREAD_UNCHANGED).astype(np.float32) * np.float32(0.0001)
        pose = np.concatenate([row[[3, 7, 11]], row[[0, 1, 2, 4, 5, 6, 8, 9, 10]]]).astype(np.float32)
        c = tuple(np.float32(BIGBIRD_CAMS[n][cam - 1]) for n in ("fx", "fy", "cx", "cy")) + (640, 480)
        if M is None:
            M = refpy.RefMap3(cam=c)
        else:
            M.set_cam(*c)
        dz = np.ascontiguousarray(D.T).ravel()          # column-major, as the mex gateway hands it over
        nz = np.flatnonzero(dz != 0).astype(np.int32)
        idx_all.append(nz); val_all.append(dz[nz]); off.append(off[-1] + len(nz))
        cams.append(cam); pose12.append(pose)
        M.update(dz, pose)
        cc, nn, tr = M.clusters()
        nleaves.append(len(cc)); nsamples.append(len(M.all_samples()))
        if count in (5, 20, 40):
            out[f"samples{count}"] = M.all_samples()
        print("seq3d frame", count, "valid", len(nz), "samples", nsamples[-1], "leaves", nleaves[-1], flush=True)
    out.update(depth_idx=np.concatenate(idx_all), depth_val=np.concatenate(val_all), depth_off=np.array(off, np.int64),
               cam=np.array(cams, np.int32), pose12=np.stack(pose12), nleaves=np.array(nleaves, np.int32),
               nsamples=np.array(nsamples, np.int32), rows=M.test(X), leaves=M.clusters()[0])
    return out


# Region of the synthetic room kept in the fixture: a patch of the +x wall. Queries live in x in [1.1, 1.72],
# y, z in [0.1, 0.7]; leaves that can be candidates of those queries have centres within 0.1 m of that box and
# their training balls reach another 0.05 m, so samples within 0.25 m are kept.
ROOM40_LO = np.array([0.85, -0.15, -0.15], np.float32)
ROOM40_HI = np.array([9.0, 0.95, 0.95], np.float32)


def room40():
    """BASELINE configs[1]: the 40 synthetic 640x480 depth frames of the box room (gpismap_b200/synth.py) mapped by
    the UNMODIFIED reference (every step of GPisMap3::update except updateGPs, which never changes a sample).
    Stored: an exact pre-order snapshot of the reference's octree around a patch of the +x wall (refpy.tree_dump:
    re-inserting the samples cannot rebuild the tree, the minimum-spacing rule would drop ~18 % of them) — what
    bench.py's reference arm and cpu_baseline load, and what the GPU pipeline must reproduce bit for bit —, the sample
    count after every frame and a SHA-256 of the complete final sample array (508,149 samples in 5,316 leaves)."""
    import hashlib
    from gpismap_b200 import synth
    M = refpy.RefMap3()
    ns = []
    for k in range(40):
        dz, pose = synth.frame(k, 40)
        M.update(dz, pose, nogp=True)
        ns.append(len(M.all_samples()))
        print("room40 frame", k, "samples", ns[-1], flush=True)
    S = M.all_samples()
    flags, smp, root = M.tree_dump(ROOM40_LO, ROOM40_HI)
    return dict(tree_flags=flags, samples=smp, root=root, lo=ROOM40_LO, hi=ROOM40_HI, nsamples=np.array(ns, np.int32),
                total=np.int64(len(S)), sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(S).tobytes()).digest(), np.uint8),
                nleaves=np.int32(len(M.clusters()[0])))


def map2d(seed):
    """2-D analogue of map3d: circle samples loaded into the reference's quadtree, every leaf trained by its own
    updateGPs, result rows of GPisMap::test_kernel (GPisMap.cpp:665-763) on random and lattice-symmetric queries."""
    rng = np.random.default_rng(seed)
    P = H.P2
    s = np.concatenate([H.circle_samples(5.3, 0.35, (1.37, -0.61), rng), H.circle_samples(2.1, 0.3, (9.13, 2.27), rng)])
    M = refpy.RefMap2()
    M.insert_samples(s)
    M.update_gps()
    centres, offs, samples, trained = H.ref_map_to_csr(M, P)
    x = np.stack([rng.uniform(-7, 13, 700), rng.uniform(-8, 7, 700)], 1).astype(np.float32)
    g = np.arange(-6.4, 12.81, 0.8)
    xs = np.stack(np.meshgrid(g, g - 3.2, indexing="ij"), -1).reshape(-1, 2).astype(np.float32)   # lattice planes, exact ties
    X = np.concatenate([x, xs])
    init = rng.uniform(size=(len(X), 6)).astype(np.float32)
    rows = M.test(X, init.copy())
    ncand = np.array([M.candidates(q, P["search"])[0].shape[0] for q in X], np.int32)
    root_c, root_half = M.root()
    return dict(samples_in=s, centres=centres, offsets=offs, samples=samples, X=X, init=init, rows=rows, ncand=ncand,
                boxes=M.cluster_boxes(), root_c=root_c, root_half=np.float32(root_half))


def map3d(seed, variant=None, rtimes=None):
    """variant / rtimes: a variant build of the reference with a params.h override (oracle/params_variants/), e.g.
    "rtimes25" with rtimes=2.5: the larger training balls of BASELINE configs[4]."""
    rng = np.random.default_rng(seed)
    P = dict(H.P3)
    if variant:
        refpy.use_variant(variant)
        P["rtimes"] = rtimes
    s = H.sphere_samples(0.12, 0.0125, (0.0517, 0.0231, 0.0113), rng)
    M = refpy.RefMap3()
    M.insert_samples(s)
    M.update_gps()
    centres, offs, samples, trained = H.ref_map_to_csr(M, P)
    d = rng.normal(size=(300, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x = (d * (0.12 + rng.uniform(-0.09, 0.09, 300))[:, None] + np.array([0.0517, 0.0231, 0.0113])).astype(np.float32)
    # lattice-symmetric queries (multiples of 0.025): exact centre-distance ties, > 16 candidates
    g = np.arange(-0.1, 0.2001, 0.025)
    xs = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    init = rng.uniform(size=(len(x) + len(xs), 8)).astype(np.float32)
    X = np.concatenate([x, xs])
    rows = M.test(X, init.copy())
    ncand = np.array([M.candidates(q, P["search"])[0].shape[0] for q in X], np.int32)
    root_c, root_half = M.root()
    out = dict(samples_in=s, centres=centres, offsets=offs, samples=samples, X=X, init=init, rows=rows, ncand=ncand,
               boxes=M.cluster_boxes(), root_c=root_c, root_half=np.float32(root_half))
    if variant:
        refpy.use_variant(None)
    return out


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:   # regenerate selected fixtures: python make_golden.py seq3d map2d seq2d_demo
        for name in only:
            fn = {"seq3d": seq3d, "map2d": lambda: map2d(15), "seq2d_demo": seq2d_demo, "room40": room40, "seq2d": seq2d,
                  "map3d_rt25": lambda: map3d(16, "rtimes25", 2.5)}[name]
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **fn())
            print(name, os.path.getsize(os.path.join(HERE, name + ".npz")))
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, "map3d_rt25.npz"), **map3d(16, "rtimes25", 2.5))
    np.savez_compressed(os.path.join(HERE, "seq3d.npz"), **seq3d())
    np.savez_compressed(os.path.join(HERE, "map2d.npz"), **map2d(15))
    np.savez_compressed(os.path.join(HERE, "leaf3d.npz"), **leaf(3, 24, 11))
    np.savez_compressed(os.path.join(HERE, "leaf2d.npz"), **leaf(2, 18, 12))
    np.savez_compressed(os.path.join(HERE, "obs2d.npz"), **obs2d(13))
    np.savez_compressed(os.path.join(HERE, "seq2d.npz"), **seq2d())
    np.savez_compressed(os.path.join(HERE, "map3d.npz"), **map3d(14))
    np.savez_compressed(os.path.join(HERE, "seq2d_demo.npz"), **seq2d_demo())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
