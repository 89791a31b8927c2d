"""Synthetic inputs of the shapes BASELINE.json names (no datasets are reachable):

* 640x480 z-depth frames of an axis-aligned box room whose walls are kept OFF the octree lattice
  (a sample with a coordinate exactly on a multiple of 0.003125 m is dropped by the reference's
  strict containsPoint), seen from a camera on a small Lissajous path whose yaw
  sweeps 2*pi over the sequence;
* the 256^3 / 512^3 query grids over the room (config 3 / 5), offset by an irrational fraction of
  the leaf pitch so no exact centre-distance ties occur.

Layouts follow the reference's entry points: depth is column-major (dataz[col*H + row],
cpp/src/GPisMap3.cpp:183), pose = [t(3) | R column-major(9)] local->global (:141-142, 201-203).
"""
import numpy as np

ROOM_LO = np.array([-1.6117, -1.3093, -1.0071])
ROOM_HI = np.array([1.6183, 1.3131, 1.0049])
CAM = dict(fx=568.0, fy=568.0, cx=310.0, cy=224.0, width=640, height=480)   # camParam defaults, GPisMap3.h:37-44


def camera_pose(k, nframes):
    """Pose k of the sequence: position on a Lissajous curve, yaw sweeping 2*pi, +-0.3 rad pitch."""
    s = 2.0 * np.pi * k / max(nframes, 1)
    t = np.array([0.30 * np.sin(2 * s + 0.3), 0.25 * np.sin(3 * s + 1.1), 0.15 * np.sin(s + 0.7)])
    yaw = s + 0.123
    pitch = 0.3 * np.sin(2 * s + 0.5)
    fwd = np.array([np.cos(yaw) * np.cos(pitch), np.sin(yaw) * np.cos(pitch), np.sin(pitch)])
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 1)   # columns: camera x (right), y (down), z (forward) in world
    return t, R


def pose12(t, R):
    return np.concatenate([t, R.T.ravel()]).astype(np.float32)   # R(:) column-major


def depth_frame(t, R, cam=CAM, lo=ROOM_LO, hi=ROOM_HI, noise_mm=0.0, rng=None):
    """z-depth image (H, W) float32 of the box room from pose (t, R); 0 = invalid."""
    H, W = cam["height"], cam["width"]
    u = (np.arange(W) - cam["cx"]) / cam["fx"]
    v = (np.arange(H) - cam["cy"]) / cam["fy"]
    uu, vv = np.meshgrid(u, v)
    d_cam = np.stack([uu, vv, np.ones_like(uu)], -1)            # z-depth parametrisation: point = z * d_cam
    d_w = d_cam @ R.T
    z = np.full((H, W), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            for plane in (lo[a], hi[a]):
                s = (plane - t[a]) / d_w[..., a]
                p = t[None, None, :] + s[..., None] * d_w
                ok = s > 0
                for b in range(3):
                    if b != a:
                        ok &= (p[..., b] >= lo[b] - 1e-9) & (p[..., b] <= hi[b] + 1e-9)
                z = np.where(ok & (s < z), s, z)
    z[~np.isfinite(z)] = 0.0
    if noise_mm > 0:
        rng = rng or np.random.default_rng(0)
        z = z + rng.normal(0.0, 1e-3 * noise_mm, z.shape) * z * z   # sigma = noise_mm mm * z^2
    return z.astype(np.float32)


def depth_colmajor(depth_hw):
    return np.ascontiguousarray(depth_hw.T).ravel()


def frame(k, nframes, noise_mm=1.0, seed=0):
    t, R = camera_pose(k, nframes)
    rng = np.random.default_rng(seed * 100003 + k)
    return depth_colmajor(depth_frame(t, R, noise_mm=noise_mm, rng=rng)), pose12(t, R)


_WALK = {}


def walk_pose(k, nframes=1000, seed=1):
    """Pose k of a bounded random-walk trajectory (BASELINE configs[3], SURVEY.md 8d: 1000 frames, seed 1): position,
    yaw and pitch do a reflected random walk inside the room; deterministic for (nframes, seed)."""
    key = (nframes, seed)
    if key not in _WALK:
        rng = np.random.default_rng(seed)
        lim = np.array([0.9, 0.7, 0.5])
        p = np.zeros(3); yaw = 0.123; pitch = 0.0
        out = []
        for _ in range(nframes):
            p = p + rng.normal(0, 0.03, 3)
            p = np.where(np.abs(p) > lim, np.sign(p) * (2 * lim - np.abs(p)), p)
            yaw += rng.normal(0.02, 0.05)
            pitch = float(np.clip(pitch + rng.normal(0, 0.03), -0.6, 0.6))
            out.append((p.copy(), yaw, pitch))
        _WALK[key] = out
    p, yaw, pitch = _WALK[key][k]
    fwd = np.array([np.cos(yaw) * np.cos(pitch), np.sin(yaw) * np.cos(pitch), np.sin(pitch)])
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    return p, np.stack([right, down, fwd], 1)


def walk_frame(k, nframes=1000, noise_mm=1.0, seed=1):
    t, R = walk_pose(k, nframes, seed)
    rng = np.random.default_rng(seed * 100003 + 7919 * k + 13)
    return depth_colmajor(depth_frame(t, R, noise_mm=noise_mm, rng=rng)), pose12(t, R)


def query_grid(n, lo=ROOM_LO, hi=ROOM_HI, inflate=0.1, z_slab=None):
    """n^3 points (or an [z0, z1) slab of them) over the inflated room box, interleaved xyz float32."""
    a = lo - inflate
    b = hi + inflate
    off = 0.05 * (np.sqrt(2.0) - 1.0) * np.array([0.31, 0.57, 0.83])   # irrational fraction of the leaf pitch
    ax = [a[c] + off[c] + (b[c] - a[c]) * (np.arange(n) + 0.5) / n for c in range(3)]
    z0, z1 = (0, n) if z_slab is None else z_slab
    zz = ax[2][z0:z1]
    X = np.empty((len(zz), n, n, 3), np.float32)
    X[..., 0] = ax[0][None, None, :]
    X[..., 1] = ax[1][None, :, None]
    X[..., 2] = zz[:, None, None]
    return X.reshape(-1, 3)
