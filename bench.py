#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: SDF + gradient + variance queries/s.

Workload (BASELINE.json configs[2], quoted on a map built from configs[1]): a 256^3 query grid
against a GPisMap3 map trained from synthetic 640x480 depth frames of the box room. One "step" is
one pass of the hot path (candidate lookup -> grouped leaf-GP evaluation -> fusion) over the grid.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference's own CPU path (oracle/_ref)

N > 1: launched by torch.distributed.run, one rank per GPU. Rank 0 builds the map through the
drop-in GPisMap3 class (host tree + GPU training), the trained leaf records are broadcast to the
other ranks over NCCL (K5), and the grid's z-planes are dealt round-robin to the ranks; the query
path itself has no collective. value = total queries / max-over-ranks device time  (strong scaling:
the grid is fixed).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="query grid edge (256 = BASELINE config 3)")
    ap.add_argument("--frames", type=int, default=40, help="synthetic depth frames used to build the map")
    ap.add_argument("--noise-mm", type=float, default=1.0)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-frames", type=int, default=2, help="--impl reference: frames the CPU reference maps itself")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation (unmodified sources in oracle/_ref) on the host cores:
    it maps `ref_frames` synthetic frames itself (GPisMap3::update) and then answers a bounded
    sample of the same query grid through GPisMap3::test with all hardware threads."""
    if rank != 0:
        return
    from gpismap_b200 import synth
    from oracle import oraclepy, refpy
    oraclepy.build()
    if not refpy.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgpisref.so was not built (needs /root/reference at build time)"}))
        return
    cores = refpy.lib().ref_hardware_concurrency()
    M = refpy.RefMap3()
    t0 = time.time()
    phases = []
    for k in range(args.ref_frames):
        dz, pose = synth.frame(k, args.frames, noise_mm=args.noise_mm)
        ph, cnt = M.update(dz, pose, timed=True)
        phases.append([round(float(x), 3) for x in ph])
    t_map = time.time() - t0
    # bounded sample of the grid: every k-th point, k chosen so one step is a few seconds of CPU
    X = synth.query_grid(args.grid)
    stride = max(1, X.shape[0] // 400_000)
    Xs = np.ascontiguousarray(X[::stride])
    times = []
    for it in range(args.warmup + args.steps):
        res = np.zeros((Xs.shape[0], 8), np.float32)
        t1 = time.perf_counter()
        M.test(Xs, res)
        dt = time.perf_counter() - t1
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    qps = Xs.shape[0] / (ms * 1e-3)
    sample = (f"every {stride}th point of the {args.grid}^3 grid ({Xs.shape[0]} queries/step) against the map the reference "
              f"built itself from {args.ref_frames} of the {args.frames} frames ({t_map:.0f} s of GPisMap3::update)")
    out = {
        "impl": "reference", "metric": "sdf_grad_var_queries_per_s", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "update_phases_s": phases,
    }
    print(json.dumps(out))


def workload_config(args):
    return {"workload": f"3D SDF query: {args.grid}^3 grid (f, grad f, variance) against a GPisMap3 map trained from "
                        f"{args.frames} synthetic 640x480 depth frames of the box room (BASELINE configs[2] on configs[1])",
            "grid": args.grid, "frames": args.frames, "depth_noise_mm_at_1m": args.noise_mm,
            "l2": "inputs larger than L2 (query + result arrays 0.7 GB, leaf records tens of GB)",
            "sharding": "z-planes round-robin over ranks, trained leaf records broadcast over NCCL"}


# ----------------------------------------------------------------------------- own arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gpismap_b200 import cabi, hostapi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ------------------------------------------------------------------ map
    t_build0 = time.time()
    update_ms = []
    train_ms = []
    gmap = None
    if rank == 0:
        gmap = hostapi.GPisMap3(device=local)
        for k in range(args.frames):
            dz, pose = synth.frame(k, args.frames, noise_mm=args.noise_mm)
            t0 = time.perf_counter()
            gmap.update(dz, pose)
            update_ms.append(1e3 * (time.perf_counter() - t0))
            train_ms.append(gmap.timing()[2])
        ctx = cabi.Ctx(3, local, borrowed=gmap.ctx_handle())
    else:
        ctx = cabi.Ctx(3, local)
    if world > 1:
        # K5: replicate the trained records. Rank 0 packs them into one device buffer; NCCL broadcast;
        # the other ranks install them in their own arena + leaf table.
        meta = torch.zeros(8, dtype=torch.int64, device=dev)
        if rank == 0:
            ptr, nbytes = ctx.export_dirty()
            rm, lv = ctx.get_rebase()
            meta[:5] = torch.tensor([nbytes, int(rm[0]), int(rm[1]), int(rm[2]), lv], dtype=torch.int64)
        dist.broadcast(meta, 0)
        nbytes = int(meta[0].item())
        CH = 1 << 30
        # view rank 0's export buffer as a torch tensor without copying
        class _Holder:
            def __init__(self, p, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 3}
        if rank == 0:
            payload = torch.as_tensor(_Holder(ptr, nbytes), device=dev)
        else:
            payload = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        for o in range(0, nbytes, CH):
            dist.broadcast(payload[o:o + CH], 0)
        if rank != 0:
            ctx.import_records(payload.data_ptr(), nbytes)
            ctx.rebase([int(meta[1]), int(meta[2]), int(meta[3])], int(meta[4]))
            del payload
        torch.cuda.synchronize()
        dist.barrier()
    t_build = time.time() - t_build0
    st0 = ctx.stats()

    # ------------------------------------------------------------------ queries: this rank's z-planes
    G = args.grid
    planes = list(range(rank, G, world))
    X = np.concatenate([synth.query_grid(G, z_slab=(z, z + 1)) for z in planes], 0)
    nq = X.shape[0]
    total_q = G ** 3
    x_host = torch.from_numpy(X).pin_memory()
    res_host = torch.zeros((nq, 8), dtype=torch.float32).pin_memory()
    x_dev = x_host.to(dev)
    res_dev = torch.zeros((nq, 8), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        res_dev.zero_()
        torch.cuda.synchronize()
        ctx.query_device(x_dev.data_ptr(), nq, res_dev.data_ptr())   # synchronous; device time from the library's CUDA events
        s = ctx.stats()
        return s["last_query_ms"], s

    for _ in range(args.warmup):
        step_device()
    sync_all()
    launches0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    sampler.start()
    ms_steps, eval_ms_steps = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ms, s = step_device()
        ms_steps.append(ms)
        eval_ms_steps.append(s["last_query_eval_ms"])
    sync_all()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0) / args.steps
    clocks = sampler.stop()
    launches = ctx.stats()["kernel_launches"] - launches0
    sq = ctx.stats()

    # end to end through the C ABI with HOST buffers: H2D of x and of res (read-modify-write), D2H of res
    e2e_ms = []
    for it in range(1 + args.steps):
        res_host.zero_()
        sync_all()
        t0 = time.perf_counter()
        rc = cabi.lib().gpis_query(ctx.h, C.c_void_p(x_host.data_ptr()), nq, C.c_void_p(res_host.data_ptr()))
        dt = time.perf_counter() - t0
        assert rc == 0
        if it > 0:
            e2e_ms.append(1e3 * dt)

    my_ms = float(np.mean(ms_steps))
    my_e2e = float(np.mean(e2e_ms))
    if world > 1:
        t = torch.tensor([my_ms, my_e2e, wall_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_all, e2e_all, wall_all = [float(v) for v in t.tolist()]
        ev = torch.tensor([float(sq["last_query_evals"])], dtype=torch.float64, device=dev)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
        evals_total = float(ev.item())
    else:
        ms_all, e2e_all, wall_all = my_ms, my_e2e, wall_ms
        evals_total = float(sq["last_query_evals"])

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_src = "fallback 6650 GB/s (B200_PROFILING.md)"
        hbm = 6650.0
        if os.path.exists(pk):
            peaks = json.load(open(pk))
            hbm = float(peaks.get("hbm_gbs", hbm))
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        eval_s = float(np.mean(eval_ms_steps)) * 1e-3
        comp_bytes = sq["last_query_bytes_compulsory"]
        achieved = comp_bytes / eval_s / 1e9 if eval_s > 0 else 0.0
        props = torch.cuda.get_device_properties(local)
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        fp32_peak = props.multi_processor_count * 128 * 2 * sm_clock / 1e12
        fp32_ach = sq["last_query_flops"] / eval_s / 1e12 if eval_s > 0 else 0.0
        # DRAM traffic of the evaluation kernels for this exact workload, from the committed ncu capture
        # (profiles/r01_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum over one step's launches)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tj.get("grid") == args.grid and tj.get("frames") == args.frames and world == 1:
                traffic = float(tj["dram_bytes_per_step"])
        except (OSError, ValueError, KeyError):
            traffic = None
        out = {
            "metric": "sdf_grad_var_queries_per_s", "value": total_q / (ms_all * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": total_q / (e2e_all * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(nq * (3 + 8) * 4), "d2h_bytes_per_step": int(nq * 8 * 4),
                    "ms_per_step": e2e_all, "note": "gpis_query() with pinned host buffers; res is read-modify-write so it is uploaded too"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "kernel": "k_eval_v3<8>/<6>/<4> (+ bucketing)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": comp_bytes, "binding": "fp32_fma (see roofline_fp32)",
                         "per": "one step = all evaluation launches of one pass over the query grid",
                         "note": "HBM term uses compulsory bytes (44 B/query + each touched leaf record once per pass); "
                                 "the binding roofline of this kernel is FP32 FMA, see roofline_fp32"},
            "roofline_fp32": {"bound": "fp32_fma", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                              "frac": fp32_ach / fp32_peak if fp32_peak else None,
                              "flops_per_step": sq["last_query_flops"],
                              "peak_source": f"{props.multi_processor_count} SMs x 128 lanes x 2 x {sm_clock/1e6:.0f} MHz (median SM clock under load)"},
            "query_breakdown": {"evaluations_per_step": evals_total, "eval_kernel_ms": float(np.mean(eval_ms_steps)),
                                "all_kernels_ms": my_ms, "wall_ms_per_step": wall_all,
                                "eval_ctas_8_6_4_1": sq["last_query_items"],
                                "batch_fill": (sq["last_query_evals"] / max(1, sum(c * q for c, q in zip(sq["last_query_items"], (8, 6, 4, 1)))))},
            "map": {"leaves": sq["leaves"], "leaves_trained": sq["leaves_trained"],
                    "arena_gb": sq["arena_bytes_used"] / 1e9, "build_s": t_build,
                    "update_ms_per_frame_median": float(np.median(update_ms)) if update_ms else None,
                    "update_ms_per_frame_p90": float(np.percentile(update_ms, 90)) if update_ms else None,
                    "train_kernel_ms_per_frame_median": float(np.median(train_ms)) if train_ms else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, gmap, X)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, gmap, X):
    """The reference's CPU path (oracle/_ref: unmodified sources + shim LA) on a bounded sample of the
    same workload: the samples of the GPU-built map inside a sub-box are loaded into the reference's
    own octree, its own updateGPs trains the leaves there, and its own test() answers the grid points
    of that sub-box with all hardware threads."""
    try:
        from oracle import oraclepy, refpy
        oraclepy.build()
        if not refpy.available():
            return {"value": None, "unit": "queries/s", "cores": None, "kind": "reference", "sample": "oracle/_ref not built"}
        from gpismap_b200 import synth
        cores = refpy.lib().ref_hardware_concurrency()
        S = gmap.all_samples()
        # sub-box: 0.6 m cube around the wall/floor edge nearest to a well-observed sample
        p0 = S[len(S) // 2, :3].astype(np.float64)
        lo = p0 - 0.3
        hi = p0 + 0.3
        m = 0.2
        sel = np.all((S[:, :3] > lo - m) & (S[:, :3] < hi + m), axis=1)
        M = refpy.RefMap3()
        M.insert_samples(S[sel])
        t0 = time.time()
        ntrain = M.update_gps(lo - 0.06, hi + 0.06)
        t_train = time.time() - t0
        q = X[np.all((X > lo) & (X < hi), axis=1)]
        # keep it to roughly the requested number of seconds
        res = np.zeros((min(len(q), 2000), 8), np.float32)
        t0 = time.perf_counter()
        M.test(q[:len(res)], res)
        rate = len(res) / (time.perf_counter() - t0)
        nq = int(min(len(q), max(2000, rate * args.cpu_baseline_seconds)))
        res = np.zeros((nq, 8), np.float32)
        t0 = time.perf_counter()
        M.test(q[:nq], res)
        dt = time.perf_counter() - t0
        ev = int((res[:, 4] < 1.0).sum())
        # the same points through the GPU map (built from the same samples): a parity report at bench-scale leaf
        # sizes, on the points whose whole candidate neighbourhood lies inside the sub-box the reference trained
        parity = None
        try:
            from gpismap_b200 import hostapi
            g2 = hostapi.GPisMap3()            # same samples, same insertion order, every leaf trained on the final set
            g2.insert_samples(S[sel])
            g2.train_active()
            got = g2.test(np.ascontiguousarray(q[:nq]))
            g2.close()
            inner = np.all((q[:nq] > lo + 0.05) & (q[:nq] < hi - 0.05), axis=1) & (res[:, 4] < 1.0) & (got[:, 4] < 1.0)
            a, b = got[inner].astype(np.float64), res[inner].astype(np.float64)
            ef = np.abs(a[:, 0] - b[:, 0]) / np.maximum(np.abs(b[:, 0]), 0.05)
            eg = np.linalg.norm(a[:, 1:4] - b[:, 1:4], axis=1) / np.maximum(np.linalg.norm(b[:, 1:4], axis=1), 0.1)
            evr = (np.abs(a[:, 4:] - b[:, 4:]) / np.maximum(np.abs(b[:, 4:]), 1e-3)).max(1)
            parity = {"rows": int(inner.sum()),
                      "f_rel": {"median": float(np.median(ef)), "p99": float(np.percentile(ef, 99)), "within_1e-4": float((ef < 1e-4).mean())},
                      "grad_rel": {"median": float(np.median(eg)), "p99": float(np.percentile(eg, 99)), "within_1e-4": float((eg < 1e-4).mean())},
                      "var_rel": {"median": float(np.median(evr)), "p99": float(np.percentile(evr, 99)), "within_1e-3": float((evr < 1e-3).mean())},
                      "note": "a second GPU map loaded with the same samples in the same order (insertSamples + trainActive) vs the reference's own train+test; floors 0.05 (f), 0.1 (|grad|), 1e-3 (var) as in tests/helpers.py; "
                              "rows the fp32 reference itself does not pin (tests/helpers.py::check_rows) are included"}
        except Exception as e:
            parity = {"error": repr(e)}
        return {"value": nq / dt, "unit": "queries/s", "cores": cores, "kind": "reference", "parity_vs_reference": parity,
                "sample": f"{nq} grid points of the sub-box {np.round(lo,2).tolist()}..{np.round(hi,2).tolist()} ({ev} evaluated) against "
                          f"{ntrain} leaves trained by the reference's updateGPs from the same samples ({t_train:.1f} s); {dt:.1f} s of GPisMap3::test",
                "leaf_train_s": t_train, "leaves_trained": ntrain}
    except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
        return {"value": None, "unit": "queries/s", "cores": None, "kind": "reference", "sample": f"failed: {e!r}"}


if __name__ == "__main__":
    main()
