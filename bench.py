#!/usr/bin/env python
"""bench.py — headline benchmark of the plane-RANSAC hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration "ms per 10M-pt multi-plane extraction" is quoted
on): a synthetic 10M-point indoor scene per GPU, 20 planes peeled iteratively, 4096 hypotheses scored per
round (max_iterations = 4095, probability = 1.0), distance threshold 0.1 (Dialog/config.txt:29), minimum
plane size 500 (Dialog/config.txt:20), PCL RNG seed 12345.  A step is one full extraction.  With N > 1
ranks the cloud is N x 10M points sharded by contiguous index range (weak scaling); per-hypothesis counts
and refit moments are summed with NCCL all-reduces, so the result equals the one-GPU result on the same
cloud bit for bit.

Printed (rank 0, one JSON line): value = point-hypotheses scored per second over the whole job with the
cloud resident in HBM; ms_per_step = ms per extraction; e2e = the same metric through the C ABI with host
buffers (pinned upload of the cloud + download of coefficients and inlier index lists inside the timed
region); roofline (scoring kernel vs the FP32-FMA peak measured live, 6 FLOP per point-hypothesis) and
roofline_hbm (compaction / refit vs MEASURED_PEAKS.json); cpu_baseline (the CPU oracle's PCL-faithful scalar
loop on a bounded sample, rank 0 only).

--impl reference times the CPU restatement of the reference's PCL path (oracle/, all host threads) on a
bounded sample of the same workload.  PCL itself is not installable here.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RANSAC point-hypotheses scored/sec; ms per 10M-pt multi-plane extraction"
UNIT = "point-hypotheses/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=10_000_000, help="points per GPU")
    ap.add_argument("--hyps", type=int, default=4096, help="hypotheses per round")
    ap.add_argument("--planes", type=int, default=20)
    ap.add_argument("--cpu-sample-hyps", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline-out", default=None,
                    help="write every rank's per-round kernel stamps of the last timed step (device-resident loop) to this JSON file")
    ap.add_argument("--scorer", default="brute", choices=["brute", "hier"],
                    help="brute = the FP32-FMA-bound kernel the roofline is reported for (default); hier = same counts "
                         "through the bounding-box culled scorer")
    ap.add_argument("--extras", action="store_true",
                    help="also time BASELINE configs[1] (1M points, 3 planes, K=1024) and a slice of configs[4] "
                         "(batch of 32K-point clouds, K=256) on rank 0's GPU; reported under 'extras'")
    ap.add_argument("--batch-clouds", type=int, default=512)
    ap.add_argument("--no-hbm-100m", action="store_true", help="skip the single-launch 100M-point compaction / refit measurement")
    ap.add_argument("--no-config4", action="store_true",
                    help="skip BASELINE configs[3]: the 100M-point scene sharded over the N ranks (strong scaling), with the "
                         "sharded result checked against one GPU solving the whole cloud")
    ap.add_argument("--config4-points", type=int, default=100_000_000)
    ap.add_argument("--no-config5", action="store_true",
                    help="skip BASELINE configs[4]: 512 x N clouds of 32K points (the full 4096-cloud batch at N = 8), sharded "
                         "by cloud id over the N ranks; also skips the configs[1] block (1M points, rank 0)")
    ap.add_argument("--config5-clouds-per-gpu", type=int, default=512)
    return ap.parse_args()


def workload_config(args, n_gpus, p2p=None):
    if n_gpus <= 1:
        sharding = "none"
    elif p2p:
        sharding = ("points, contiguous index ranges; per round the sample points, int32 counts, int64 moments and remaining "
                    "counts are exchanged by single-kernel peer-memory exchanges over NVLink (CUDA IPC mailboxes)")
    else:
        sharding = "points, contiguous index ranges; NCCL all-reduce of int32 counts + int64 moments"
    return {
        "workload": "configs[2]: synthetic indoor scene, %d points per GPU x %d GPU(s), %d planes peeled, %d "
                    "hypotheses per round (max_iterations=%d, probability=1.0), t=0.1, min_plane=500, seed=12345"
                    % (args.points, n_gpus, args.planes, args.hyps, args.hyps - 1),
        "points_per_gpu": args.points, "points_total": args.points * n_gpus, "hypotheses_per_round": args.hyps,
        "planes": args.planes, "distance_threshold": 0.1, "min_plane_size": 500, "dot_order": "fma", "scorer": args.scorer,
        "sharding": sharding,
        "l2": "256 MiB fill kernel between timed steps (outside the timed region); cloud planes are 120 MB per GPU",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes per launch of the scoring kernel from the committed ncu capture (profiles/r02_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def cpu_sample(args, pts, threads):
    """PCL's countWithinDistance loop (the oracle) over the first S hypotheses of round 0."""
    from oracle import oracle as O
    S = args.cpu_sample_hyps
    tri = O.draw_sequence(pts.shape[0], S)
    coeffs, good = O.models_from_triples(pts, tri)
    t0 = time.perf_counter()
    O.count_batch(pts, np.nan_to_num(coeffs), 0.1, O.DOT_FMA, threads=threads)
    dt = time.perf_counter() - t0
    return pts.shape[0] * S / dt, dt


def run_reference(args, out):
    """Reference arm: the CPU restatement of the PCL path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dialog_b200 import synth
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1; this arm uses every host thread
    pts = synth.indoor_scene().points(0, args.points)
    from oracle import oracle as O
    S = max(8, min(args.cpu_sample_hyps, 64))
    tri = O.draw_sequence(pts.shape[0], S)
    coeffs, good = O.models_from_triples(pts, tri)
    coeffs = np.nan_to_num(coeffs)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.count_batch(pts, coeffs, 0.1, O.DOT_FMA, threads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = pts.shape[0] * S / (ms * 1e-3)
    sample = "countWithinDistance of the first %d hypotheses of round 0 over the %d-point cloud per step (the full " \
             "step scores %d x %d rounds); OpenMP over points, %d threads" % (S, args.points, args.hyps, args.planes, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "PCL 1.8 is not installable here; this is the CPU oracle (oracle/pr_oracle.c) restating its loop"}
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_extras(args, pr):
    """The other single-GPU BASELINE configs, resident timing, rank 0 only."""
    import dialog_b200 as D
    from dialog_b200 import synth
    out = {}
    # configs[1]: 1M points, 3 planes + 1 % noise + 30 % outliers, 1024 hypotheses
    pts = synth.three_planes_scene().points(0, 1_000_000)
    pr.set_cloud(pts)
    prm = D.make_params(0.1, 1023, 500, 1.0, True, 12345, 3, D.DOT_FMA)
    for _ in range(3):
        ex = pr.extract_planes(prm, want_indices=False)
    ms = []
    for _ in range(5):
        pr.flush_l2()
        pr.timer_start()
        ex = pr.extract_planes(prm, want_indices=False)
        ms.append(pr.timer_stop())
    pairs = sum(int(i.n_cloud) * int(i.n_scored) for i in ex.infos)
    out["configs[1]_1M_3planes_K1024"] = {"ms_per_extraction": sum(ms) / len(ms), "planes": len(ex.planes),
                                          "point_hypotheses_per_s": pairs / (sum(ms) / len(ms) * 1e-3)}
    # BASELINE.md plan (a): PCL-default adaptive mode (max_iterations=50, probability=0.99) end to end on the 10M-point
    # scene, GPU through the C ABI vs the CPU oracle (1 thread, the PCL-faithful loop), same planes required
    from oracle import oracle as O
    pts10 = synth.indoor_scene().points(0, args.points)
    prm = D.make_params(0.1, 50, 500, 0.99, True, 12345, args.planes, D.DOT_FMA)
    pin10 = D.PinnedArray(pts10.shape, np.float32)   # the caller's cloud in page-locked memory
    pin10.array[:] = pts10
    pr.set_cloud(pts10)
    for _ in range(2):
        exd = pr.extract_planes(prm, want_indices=True, copy=False)
    ms = []
    for _ in range(3):
        pr.flush_l2()
        pr.timer_start()
        pr.set_cloud_ptr(pin10.ptr, pts10.shape[0], overlap=True)
        exd = pr.extract_planes(prm, want_indices=True, copy=False)
        ms.append(pr.timer_stop())
    pin10.free()
    t0 = time.perf_counter()
    want = O.extract_planes(pts10, O.make_params(0.1, 50, 500, 0.99, True, 12345, args.planes, O.DOT_FMA, O.REFIT_FIXED))
    cpu_s = time.perf_counter() - t0
    same = len(want.coeffs) == len(exd.planes) and all(
        p.coeff.tobytes() == want.coeffs[k].tobytes() and np.array_equal(p.inliers_orig, want.inliers_orig[k])
        for k, p in enumerate(exd.planes))
    out["pcl_default_adaptive_10M_end_to_end"] = {
        "gpu_ms": sum(ms) / len(ms), "cpu_oracle_s": cpu_s, "cpu_threads": 1, "planes": len(exd.planes),
        "identical_to_cpu_oracle": bool(same), "speedup": cpu_s / (sum(ms) / len(ms) * 1e-3),
        "note": "pinned host cloud in (chunked upload overlapped with the first scoring pass), coefficients + inlier indices "
                "out; max_iterations=50, probability=0.99"}
    # the reference's postProcessPlanes re-absorption pass over what the extraction left, against the
    # scene's patch outlines as the plane polygons (180 vertices each); CPU = the oracle restating isPointInPoly
    # (identical to the reference's own source, tests/test_reabsorb.py), timed on a prefix of the remaining cloud
    # (extraction at half the threshold, so that plane points beyond it are left for the pass to claim at T = 0.1)
    scene = synth.indoor_scene()
    pr.set_cloud(pts10)
    ext = pr.extract_planes(D.make_params(0.05, 50, 500, 0.99, True, 12345, args.planes, D.DOT_FMA), want_indices=False)
    rem = pr.remaining().copy()
    coeffs = np.array([p.coeff for p in ext.planes], np.float32)
    # polygon of plane k = outline of the generating patch whose normal / offset it matches best
    borders = []
    for c in coeffs:
        err = [min(np.abs(q.coeff - c).max(), np.abs(q.coeff + c).max()) for q in scene.patches]
        borders.append(scene.patches[int(np.argmin(err))].border())
    ms = []
    for rep in range(4):
        pr.set_cloud(rem)
        pr.flush_l2()
        pr.timer_start()
        cur, _, n_left = pr.reabsorb(coeffs, borders, 0.1, 20261018)
        t_ms = pr.timer_stop()
        if rep:  # the first call sizes the scratch buffers
            ms.append(t_ms)
    n_cpu = min(len(rem), 20000)
    t0 = time.perf_counter()
    want = O.reabsorb(rem[:n_cpu], coeffs, borders, 0.1, 20261018)
    cpu_s = time.perf_counter() - t0
    same = all(np.array_equal(a[a < n_cpu], b) for a, b in zip(cur, want.absorbed))
    # the same prefix through the reference's OWN isPointInPoly source (oracle/_ref, built by oracle/build_ref.py)
    ref_s, ref_same = None, None
    if O.ref_lib() is not None:
        t0 = time.perf_counter()
        ref_masks = [O.ref_points_in_poly(rem[:n_cpu], coeffs[k], borders[k], 0.1, 20261018) for k in range(len(coeffs))]
        ref_s = time.perf_counter() - t0
        ref_same = all(np.array_equal(np.nonzero(m)[0], a[a < n_cpu]) for m, a in zip(ref_masks, cur))
    out["reabsorb_postProcessPlanes"] = {
        "points": int(len(rem)), "planes": int(len(coeffs)), "border_vertices_per_plane": int(len(borders[0])),
        "gpu_ms": sum(ms) / len(ms), "absorbed": int(sum(len(a) for a in cur)), "points_left": int(n_left),
        "cpu_oracle_s_per_point": cpu_s / n_cpu, "cpu_sample_points": n_cpu, "cpu_threads": 1,
        "cpu_s_extrapolated": cpu_s / n_cpu * len(rem), "identical_on_cpu_sample": bool(same),
        "speedup_extrapolated": (cpu_s / n_cpu * len(rem)) / (sum(ms) / len(ms) * 1e-3),
        "reference_source": None if ref_s is None else {
            "kind": "reference", "what": "Dialog/PlaneDetect.h isPointInPoly compiled over oracle/ref_shim.h (oracle/_ref), 1 thread",
            "s_per_point": ref_s / n_cpu, "s_extrapolated": ref_s / n_cpu * len(rem), "identical_on_cpu_sample": bool(ref_same),
            "speedup_extrapolated": (ref_s / n_cpu * len(rem)) / (sum(ms) / len(ms) * 1e-3)},
        "note": "cloud resident (set_cloud before the timer), polygon upload + kernels + index lists back inside it"}
    # estimateNormal() (pcl::NormalEstimationOMP, radius search) on the 10M-point scene
    pr.set_cloud(pts10)
    pr.estimate_normals(0.1)
    ms = []
    for _ in range(2):
        pr.timer_start()
        nrm, ncnt = pr.estimate_normals(0.1, want_counts=True)
        ms.append(pr.timer_stop())
    n_chk = 4000
    want_n, want_c = O.estimate_normals(pts10[:n_chk], 0.1)   # brute-force oracle on a prefix: neighbourhoods differ from
    sub_n, sub_c = None, None                                  # the full cloud's, so the check stages the same prefix
    pr.set_cloud(pts10[:n_chk])
    sub_n, sub_c = pr.estimate_normals(0.1, want_counts=True)
    okn = want_c >= 3
    dn = np.minimum(np.abs(sub_n[okn][:, :3] - want_n[okn][:, :3]).max(1), np.abs(sub_n[okn][:, :3] + want_n[okn][:, :3]).max(1))
    out["estimate_normals_10M_r0.1"] = {
        "gpu_ms": min(ms), "mean_neighbours": float(ncnt.mean()), "neighbour_pairs_per_s": float(ncnt.astype(np.int64).sum()) / (min(ms) * 1e-3),
        "nan_normals": int(np.isnan(nrm[:, 0]).sum()),
        "oracle_check": {"points": n_chk, "counts_identical": bool(np.array_equal(sub_c, want_c)), "max_normal_diff": float(dn.max()) if okn.any() else 0.0},
        "note": "includes the 160 MB + 40 MB download of normals and counts to pageable host memory"}
    # BASELINE configs[3] scale, PCL-default adaptive mode (max_iterations = 50, probability = 0.99), host cloud in ->
    # coefficients + inlier indices out, against the CPU oracle on the whole 100M-point scene (~100 s of one core)
    big = storeys_cloud(pts10, 0, 100_000_000)
    pinb = D.PinnedArray(big.shape, np.float32)
    pinb.array[:] = big
    prm = D.make_params(0.1, 50, 500, 0.99, True, 12345, args.planes, D.DOT_FMA)
    pr.set_cloud_ptr(pinb.ptr, big.shape[0], overlap=True)
    exb = pr.extract_planes(prm, want_indices=True, copy=False)
    ms = []
    for _ in range(2):
        pr.flush_l2()
        pr.timer_start()
        pr.set_cloud_ptr(pinb.ptr, big.shape[0], overlap=True)
        exb = pr.extract_planes(prm, want_indices=True, copy=False)
        ms.append(pr.timer_stop())
    t0 = time.perf_counter()
    want = O.extract_planes(big, O.make_params(0.1, 50, 500, 0.99, True, 12345, args.planes, O.DOT_FMA, O.REFIT_FIXED))
    cpu_s = time.perf_counter() - t0
    same = len(want.coeffs) == len(exb.planes) and all(
        p.coeff.tobytes() == want.coeffs[k].tobytes() and np.array_equal(p.inliers_orig, want.inliers_orig[k])
        for k, p in enumerate(exb.planes))
    out["pcl_default_adaptive_100M_end_to_end"] = {
        "gpu_ms": sum(ms) / len(ms), "cpu_oracle_s": cpu_s, "cpu_threads": 1, "planes": len(exb.planes),
        "identical_to_cpu_oracle": bool(same), "speedup": cpu_s / (sum(ms) / len(ms) * 1e-3),
        "note": "ten 10M-point storeys (the configs[3] scene); 1.6 GB pinned cloud in, coefficients + inlier indices out"}
    pinb.free()
    del big, want
    pr.set_cloud(pts10[:1000])
    # the drop-in C++ surface end to end at the headline size: examples/plane_detect_demo --bench (PlaneDetectRansac::
    # detectViews on a page-locked cloud: upload + 20 rounds + index lists back), wall clock per call
    try:
        import tempfile
        from dialog_b200 import build as _b
        demo = _b.build_demo()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "cloud.f32")
            pts10[:, :3].astype("<f4").tofile(path)
            r = subprocess.run([demo, path, "0.1", str(args.hyps - 1), "500", "--prob", "1.0", "--max-planes", str(args.planes), "--bench", "5"],
                               capture_output=True, text=True, timeout=600)
        f = r.stdout.split()
        out["cpp_shim_end_to_end"] = {"ms_per_step": float(f[f.index("e2e_ms_per_step") + 1]), "planes": int(f[f.index("planes") + 1]),
                                      "inliers": int(f[f.index("inliers") + 1]),
                                      "what": "examples/plane_detect_demo --bench 5: PlaneDetectRansac::detectViews (include/PlaneDetectRansac.h) on a "
                                              "page-locked 10M-point cloud, same parameters as the headline; compare with e2e.ms_per_step"}
    except Exception as e:  # noqa: BLE001
        out["cpp_shim_end_to_end"] = {"error": str(e)}
    # configs[4]: batch of 32K-point clouds, one plane each, 256 hypotheses per cloud (slice of the 4096 clouds)
    nc = args.batch_clouds
    clouds = np.stack([synth.tile_scene(cid).points(0, 32768) for cid in range(nc)])
    pr.set_cloud_batch(clouds)
    prm = D.make_params(0.1, 255, 500, 1.0, True, 12345, 1, D.DOT_FMA)
    for _ in range(2):
        pr.segment_batch(prm, want_infos=False)
    ms = []
    for _ in range(5):
        pr.flush_l2()
        pr.timer_start()
        coeffs, cnt, _ = pr.segment_batch(prm, want_infos=False)
        ms.append(pr.timer_stop())
    pairs = nc * 32768 * 257
    out["configs[4]_batch_32K_clouds_K256"] = {"clouds": nc, "ms_per_batch": sum(ms) / len(ms),
                                               "clouds_per_s": nc / (sum(ms) / len(ms) * 1e-3),
                                               "point_hypotheses_per_s": pairs / (sum(ms) / len(ms) * 1e-3),
                                               "mean_inliers": float(cnt.mean())}
    return out


def _mix64(g):
    """Order-independent 64-bit hash ingredient of an index array (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (g.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
        return x


def storeys_cloud(base, first, count):
    """BASELINE configs[3] scene: global point i is point (i mod len(base)) of the 10M-point indoor storey, lifted by
    8 m per storey (i div len(base)); rows [first, first + count)."""
    nb = base.shape[0]
    out = np.empty((count, 4), np.float32)
    done = 0
    while done < count:
        g = first + done
        storey, off = divmod(g, nb)
        take = min(nb - off, count - done)
        out[done:done + take] = base[off:off + take]
        out[done:done + take, 2] += np.float32(8.0 * storey)
        done += take
    return out


def run_config4(args, torch, dist, D, local_rank, world, rank, hbm_peak, peak_tf):
    """BASELINE configs[3]: one 100M-point scene, sharded by contiguous index range over the N ranks (strong scaling),
    20 planes x 4096 hypotheses.  Outside the headline timer.  With N > 1 rank 0 also solves the whole cloud on its own
    GPU (the shards are gathered over NCCL) and the sharded planes must equal it: coefficients, global inlier counts and
    an order-independent hash of every plane's global inlier indices."""
    from dialog_b200 import synth
    n_total = args.config4_points
    base = synth.indoor_scene().points(0, min(10_000_000, n_total))
    first, count = D.host_shard_range(n_total, world, rank)
    shard = storeys_cloud(base, first, count)
    del base
    pinned = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = shard
    del shard
    prm = D.make_params(0.1, args.hyps - 1, 500, 1.0, True, 12345, args.planes, D.DOT_FMA)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pr = D.PlaneRansac(local_rank)
    if world > 1:
        uid = [D.PlaneRansac.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pr.comm_init(world, rank, uid[0])
    pr.set_cloud_ptr(pinned.data_ptr(), count)
    ex = pr.extract_planes(prm, want_indices=True, copy=False)   # warm-up; also the result that is checked
    pr.profile_enable(True)
    pr.profile_reset()
    ms = []
    for _ in range(2):
        pr.flush_l2()
        barrier()
        pr.timer_start()
        ex = pr.extract_planes(prm, want_indices=False)
        ms.append(pr.timer_stop())
    barrier()
    prof = pr.profile()
    pr.profile_enable(False)
    ex = pr.extract_planes(prm, want_indices=True, copy=False)
    pairs = sum(int(i.n_cloud) * int(i.n_scored) for i in ex.infos)
    # per-plane fingerprints of this rank's part, summed over ranks
    P = len(ex.planes)
    fp = np.zeros((args.planes, 2), np.int64)
    for k, p in enumerate(ex.planes):
        g = p.inliers_orig.astype(np.int64) + first
        fp[k, 0] = p.inliers_orig.size
        fp[k, 1] = _mix64(g).sum(dtype=np.uint64).astype(np.int64) if g.size else 0
    t = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
    fpt = torch.from_numpy(fp).cuda()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fpt, op=dist.ReduceOp.SUM)
    ms_step = t.item() / 2
    out = {"points_total": n_total, "n_gpus": world, "planes": P, "hypotheses_per_round": args.hyps,
           "scaling": "strong", "ms_per_extraction": ms_step, "value": pairs / (ms_step * 1e-3), "unit": UNIT,
           "scene": "ten 10M-point indoor storeys 8 m apart (walls coplanar across storeys), contiguous index-range shards"}
    if rank == 0:
        score_tf = 6.0 * prof.pairs_scored / (prof.ms_score * 1e-3) / 1e12 if prof.ms_score > 0 else None
        out["roofline"] = {"kernel": "score_kernel<8,FMA>", "bound": "fp32_fma", "achieved": score_tf, "peak": peak_tf, "unit": "TFLOP/s",
                           "frac": score_tf / peak_tf if score_tf else None, "note": "rank 0's launches"}
        out["roofline_hbm"] = hbm_block(prof, hbm_peak)
        out["kernel_ms_per_extraction"] = {"models_draw": prof.ms_models / 2, "score": prof.ms_score / 2, "refit": prof.ms_refit / 2,
                                           "compact": prof.ms_compact / 2, "other": prof.ms_other / 2}
    # one GPU on the whole cloud (rank 0): gathered shards -> plane_ransac_set_cloud_device
    if world > 1:
        dev = pinned.cuda()
        sizes = [D.host_shard_range(n_total, world, r)[1] for r in range(world)]
        if rank == 0:
            whole_t = torch.empty((n_total, 4), dtype=torch.float32, device="cuda")
            parts = list(whole_t.split(sizes))
            dist.gather(dev, parts, dst=0)
        else:
            dist.gather(dev, None, dst=0)
        del dev
        if rank == 0:
            one = D.PlaneRansac(local_rank)
            one.set_cloud_device_ptr(whole_t.data_ptr(), n_total)
            del whole_t, parts
            one.extract_planes(prm, want_indices=False)
            one_ms = []
            for _ in range(2):
                one.flush_l2()
                one.timer_start()
                ex1 = one.extract_planes(prm, want_indices=False)
                one_ms.append(one.timer_stop())
            ex1 = one.extract_planes(prm, want_indices=True, copy=False)
            same = len(ex1.planes) == P
            got = fpt.cpu().numpy()
            for k, p in enumerate(ex1.planes):
                if not same:
                    break
                h = _mix64(p.inliers_orig.astype(np.int64)).sum(dtype=np.uint64).astype(np.int64) if p.inliers_orig.size else 0
                same = (p.coeff.tobytes() == ex.planes[k].coeff.tobytes() and p.info.n_inliers == ex.planes[k].info.n_inliers
                        and int(got[k, 0]) == p.inliers_orig.size and int(got[k, 1]) == int(h))
            one.close()
            t1 = sum(one_ms) / 2
            out["one_gpu_ms_per_extraction"] = t1
            out["efficiency_vs_1gpu"] = t1 / (world * ms_step)
            out["sharded_identical_to_single_gpu"] = bool(same)
            out["identity_check"] = "coefficients bit for bit, global inlier count and an order-independent 64-bit hash of the global inlier indices, every plane"
        dist.barrier()
    else:
        out["efficiency_vs_1gpu"] = 1.0
        out["sharded_identical_to_single_gpu"] = None
        if not args.no_hbm_100m:
            # one segment() + peel over the whole staged cloud (round 0: 12 B per point read), three times
            prm1 = D.make_params(0.1, 255, 500, 1.0, True, 12345, 1, D.DOT_FMA)
            pr.extract_planes(prm1, want_indices=False)
            pr.profile_enable(True)
            pr.profile_reset()
            for _ in range(3):
                pr.extract_planes(prm1, want_indices=False)
            pb = pr.profile()
            pr.profile_enable(False)
            hb = hbm_block(pb, hbm_peak)
            hb["compact"]["ms"] = pb.ms_compact / 3
            hb["refit"]["ms"] = pb.ms_refit / 3
            hb["compact"]["bytes"] = int(pb.bytes_compact // 3)
            out["single_launch"] = hb
    pr.close()
    return out if rank == 0 else None


def run_config5(args, torch, dist, D, local_rank, world, rank, peak_tf):
    """BASELINE configs[4]: a batch of 32K-point clouds (per-scan tiles), one segment() each with 256 hypotheses, sharded
    by cloud id (cloud_id % N == rank) — replicas only, no data-path collective; one gather of the results at the end.
    512 clouds per GPU: the full 4096-cloud batch at N = 8, a slice of it below."""
    from dialog_b200 import synth
    n_per, K = 32768, 256
    total = args.config5_clouds_per_gpu * world
    ids = [cid for cid in range(total) if cid % world == rank]
    pinned = torch.empty((len(ids), n_per, 4), dtype=torch.float32, pin_memory=True)
    host = pinned.numpy()
    for j, cid in enumerate(ids):
        host[j] = synth.tile_scene(cid).points(0, n_per)
    prm = D.make_params(0.1, K - 1, 500, 1.0, True, 12345, 1, D.DOT_FMA)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pr = D.PlaneRansac(local_rank)
    pr.set_cloud_batch_ptr(pinned.data_ptr(), len(ids), n_per)
    for _ in range(3):
        coeffs, cnt, _ = pr.segment_batch(prm, want_infos=False)
    ms, e2e = [], []
    for _ in range(5):      # timed without events between the launches (they chain with programmatic dependent launch)
        pr.flush_l2()
        barrier()
        pr.timer_start()
        coeffs, cnt, _ = pr.segment_batch(prm, want_infos=False)
        ms.append(pr.timer_stop())
    pr.profile_enable(True)
    pr.profile_reset()
    ev = []
    for _ in range(5):      # the same batches with CUDA events around every kernel class, for the breakdown
        pr.flush_l2()
        barrier()
        pr.timer_start()
        pr.segment_batch(prm, want_infos=False)
        ev.append(pr.timer_stop())
    prof = pr.profile()
    pr.profile_enable(False)
    # host tiles in -> coefficients, counts and every cloud's inlier index list out (page-locked buffers both ways)
    lists_pin = D.PinnedArray((len(ids) * n_per,), np.int32)
    for rep in range(5):
        pr.flush_l2()
        barrier()
        pr.timer_start()
        pr.set_cloud_batch_ptr(pinned.data_ptr(), len(ids), n_per)
        coeffs, cnt, _, lists = pr.segment_batch(prm, want_infos=False, want_lists=True, lists_buf=lists_pin.array)
        if rep >= 2:
            e2e.append(pr.timer_stop())
        else:
            pr.timer_stop()
    barrier()
    t = torch.tensor([sum(ms) / len(ms), sum(e2e) / len(e2e)], dtype=torch.float64, device="cuda")
    res = torch.from_numpy(np.concatenate([coeffs, cnt[:, None].astype(np.float32)], 1)).cuda()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parts = [torch.empty_like(res) for _ in range(world)] if rank == 0 else None
        dist.gather(res, parts, dst=0)       # the one exchange of the config: results to rank 0
        if rank == 0:
            res = torch.cat(parts)
    pr.close()
    if rank != 0:
        return None
    ms_b, ms_e = t.tolist()
    pairs = total * n_per * K
    score_tf = 6.0 * prof.pairs_scored / (prof.ms_score * 1e-3) / 1e12 if prof.ms_score > 0 else None
    return {"clouds_total": total, "clouds_per_gpu": len(ids), "points_per_cloud": n_per, "hypotheses_per_cloud": K, "n_gpus": world,
            "sharding": "cloud_id % n_gpus, no collective; one gather of (coefficients, count) per cloud at the end",
            "ms_per_batch": ms_b, "clouds_per_s": total / (ms_b * 1e-3), "point_hypotheses_per_s": pairs / (ms_b * 1e-3),
            "frac_of_fp32_peak_whole_call": 6.0 * pairs / (ms_b * 1e-3) / 1e12 / (peak_tf * world),
            "score_kernel_frac_of_fp32_peak": score_tf / peak_tf if score_tf else None,
            "kernel_ms": {"gather_models": prof.ms_models / 5, "score": prof.ms_score / 5, "refit": prof.ms_refit / 5,
                          "final_count": prof.ms_compact / 5, "other": prof.ms_other / 5,
                          "event_pass_ms_per_batch": sum(ev) / len(ev),
                          "note": "separate pass with CUDA events around every kernel class (about 0.07 ms slower per batch)"},
            "e2e_ms_per_batch": ms_e, "e2e_clouds_per_s": total / (ms_e * 1e-3),
            "e2e_note": "pinned host tiles in (%d MB per GPU), coefficients + counts + every cloud's inlier index list out into "
                        "a page-locked buffer; 2 warm-up batches, 3 timed" % (len(ids) * n_per * 16 // 2**20),
            "mean_inliers": float(res[:, 4].mean().item()), "clouds_with_plane": int((res[:, 4] >= 500).sum().item())}


def run_config2(D, local_rank, peak_tf):
    """BASELINE configs[1]: 1M points, 3 planes + 1 % noise + 30 % outliers, 1024 hypotheses per round, one GPU (rank 0)."""
    from dialog_b200 import synth
    pts = synth.three_planes_scene().points(0, 1_000_000)
    pr = D.PlaneRansac(local_rank)
    pr.set_cloud(pts)
    prm = D.make_params(0.1, 1023, 500, 1.0, True, 12345, 3, D.DOT_FMA)
    for _ in range(3):
        pr.extract_planes(prm, want_indices=False)
    pr.profile_reset()
    ms = []
    for _ in range(5):
        pr.flush_l2()
        pr.timer_start()
        ex = pr.extract_planes(prm, want_indices=False)
        ms.append(pr.timer_stop())
    prof = pr.profile()
    pr.close()
    pairs = sum(int(i.n_cloud) * int(i.n_scored) for i in ex.infos)
    step = sum(ms) / len(ms)
    loop = list(prof.loop_ms)
    return {"points": 1_000_000, "planes": len(ex.planes), "hypotheses_per_round": 1024, "ms_per_extraction": step,
            "point_hypotheses_per_s": pairs / (step * 1e-3),
            "frac_of_fp32_peak_whole_call": 6.0 * pairs / (step * 1e-3) / 1e12 / peak_tf,
            "score_kernel_frac_of_fp32_peak": 6.0 * prof.pairs_scored / (loop[3] * 1e-3) / 1e12 / peak_tf if loop[3] > 0 else None,
            "device_loop_ms": dict({nm: loop[i] / len(ms) for i, nm in enumerate(D.LOOP_STAGE_NAMES) if loop[i] > 0},
                                   rounds=prof.loop_rounds / len(ms), step_minus_rounds=step - sum(loop) / len(ms))}


def hbm_block(prof, hbm_peak):
    """HBM rooflines of the peel (K5) and refit (K3) launches in a profile: bytes actually moved, and SURVEY §8d's formula."""
    c_gbs = prof.bytes_compact / (prof.ms_compact * 1e-3) / 1e9 if prof.ms_compact > 0 else None
    survey = 16 * prof.points_compact + 16 * prof.points_kept + 4 * prof.points_peeled
    s_gbs = survey / (prof.ms_compact * 1e-3) / 1e9 if prof.ms_compact > 0 else None
    r_gbs = prof.bytes_refit / (prof.ms_refit * 1e-3) / 1e9 if prof.ms_refit > 0 else None
    return {
        "compact": {"bound": "hbm", "achieved": c_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": c_gbs / hbm_peak if c_gbs else None,
                    "ms": prof.ms_compact, "bytes": int(prof.bytes_compact),
                    "algorithmic": "bytes moved: 12 B per point read (16 B once the cloud carries its original-index plane, "
                                   "i.e. after the first peel) + 16 B per kept point + 4 B per inlier per list written",
                    "frac_survey_8d_formula": s_gbs / hbm_peak if s_gbs else None,
                    "survey_8d_formula": "16 N + 16 N_rem + 4 N_inl (charges an index-plane read the staged cloud does not have)"},
        "refit": {"bound": "hbm", "achieved": r_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": r_gbs / hbm_peak if r_gbs else None,
                  "ms": prof.ms_refit, "algorithmic": "12 B read per point"}}


def claim_stdout():
    """Keep fd 1 for the JSON line only: libraries (NCCL prints its version banner) write to stderr instead."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse_args()
    out = claim_stdout()
    if args.impl == "reference":
        run_reference(args, out)
        return

    import torch
    import torch.distributed as dist
    import dialog_b200 as D
    from dialog_b200 import build, synth
    build.build()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the plane-RANSAC backend has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- data: this rank's contiguous shard of the global cloud, generated in place ----
    n_total = args.points * world
    first, count = D.host_shard_range(n_total, world, rank)
    scene = synth.indoor_scene()
    pts = scene.points(first, first + count)
    pinned = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = pts

    pr = D.PlaneRansac(local_rank)
    if world > 1:
        uid = [D.PlaneRansac.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pr.comm_init(world, rank, uid[0])
    prm = D.make_params(0.1, args.hyps - 1, 500, 1.0, True, 12345, args.planes, D.DOT_FMA,
                        D.SCORER_HIER if args.scorer == "hier" else D.SCORER_BRUTE)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pairs_of(ex):
        return sum(int(i.n_cloud) * int(i.n_scored) for i in ex.infos)

    # ---- resident arm: cloud staged once, timed region = extract_planes (coefficients only) ----
    pr.set_cloud_ptr(pinned.data_ptr(), count)
    peak_tf = pr.measure_ffma_peak()
    for _ in range(args.warmup):
        ex = pr.extract_planes(prm, want_indices=False)
    # No events between the kernels of the timed steps (an event between two kernels of a round would keep the second
    # from being set up while the first drains): the kernels of the device-resident loop stamp %globaltimer themselves
    # (pr_profile.loop_ms), which is where the roofline's kernel time comes from.
    pr.profile_reset()
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    step_ms = []
    for _ in range(args.steps):
        pr.flush_l2()
        barrier()
        pr.timer_start()
        ex = pr.extract_planes(prm, want_indices=False)
        step_ms.append(pr.timer_stop())
    barrier()
    prof_t = pr.profile()
    timeline = pr.round_timeline()
    pairs_step = pairs_of(ex)
    n_planes = len(ex.planes)
    total_ms = sum(step_ms)
    # the same steps once more with CUDA events around every kernel class (outside the timed region: the events cost
    # about 1 ms per step): per-class kernel times, HBM rooflines, exchange waits
    pr.profile_enable(True)
    pr.profile_reset()
    ev_ms = []
    for _ in range(args.steps):
        pr.flush_l2()
        barrier()
        pr.timer_start()
        pr.extract_planes(prm, want_indices=False)
        ev_ms.append(pr.timer_stop())
    barrier()
    prof = pr.profile()
    pr.profile_enable(False)

    # ---- end-to-end arm: host cloud in, coefficients + inlier index lists out, every step ----
    for _ in range(min(args.warmup, 2)):
        pr.set_cloud_ptr(pinned.data_ptr(), count, overlap=True)
        pr.extract_planes(prm, want_indices=True, copy=False)
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        pr.flush_l2()
        barrier()
        pr.timer_start()
        pr.set_cloud_ptr(pinned.data_ptr(), count, overlap=True)      # chunked upload, scored as the chunks land
        ex2 = pr.extract_planes(prm, want_indices=True, copy=False)   # index lists land in pinned host buffers
        e2e_ms.append(pr.timer_stop())
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    e2e_total_ms = sum(e2e_ms)
    n_inl_local = sum(p.inliers_cur.size for p in ex2.planes)
    n_draws = sum(int(i.n_scored) for i in ex2.infos)
    h2d = count * 16 + n_draws * 12
    d2h = n_inl_local * 8 + n_draws * 8 + len(ex2.infos) * (144 + 16 + 16)

    # ---- where the end-to-end overhead goes: the upload alone (all ranks at once: they share the host's memory system and
    #      PCIe root complexes), and the same end-to-end steps double-buffered over two contexts, so that cloud s + 1 travels
    #      while cloud s is being extracted (the ingestion pattern of a scanner feeding clouds one after the other) ----
    up_ms = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        pr.set_cloud_ptr(pinned.data_ptr(), count)          # synchronous: copy + staging kernel
        torch.cuda.synchronize()
        up_ms.append((time.perf_counter() - t0) * 1e3)
    pr2 = D.PlaneRansac(local_rank)
    if world > 1:
        uid2 = [D.PlaneRansac.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid2, src=0)
        pr2.comm_init(world, rank, uid2[0])
    ctxs = [pr, pr2]
    for cx in ctxs:
        cx.set_cloud_ptr(pinned.data_ptr(), count, overlap=True)
        cx.extract_planes(prm, want_indices=True, copy=False)
    pr.set_cloud_ptr(pinned.data_ptr(), count, overlap=True)   # step 0's cloud is on its way when the clock starts ...
    barrier()
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        ctxs[(s_ + 1) % 2].set_cloud_ptr(pinned.data_ptr(), count, overlap=True)   # ... and every step queues the next one's upload
        ctxs[s_ % 2].extract_planes(prm, want_indices=True, copy=False)
    barrier()
    db_total_ms = (time.perf_counter() - t0) * 1e3
    pr2.close()
    pr.set_cloud_ptr(pinned.data_ptr(), count)

    # ---- the PCL-faithful dot order (Eigen's SSE2 reduction: 3 FMUL + 3 FADD, separately rounded) on the same workload ----
    sse2 = None
    if args.scorer == "brute":
        prm_s = D.make_params(0.1, args.hyps - 1, 500, 1.0, True, 12345, args.planes, D.DOT_PCL_SSE2)
        for _ in range(2):
            exs = pr.extract_planes(prm_s, want_indices=False)
        pr.profile_enable(True)
        pr.profile_reset()
        sse2_ms = []
        for _ in range(min(args.steps, 3)):
            pr.flush_l2()
            barrier()
            pr.timer_start()
            exs = pr.extract_planes(prm_s, want_indices=False)
            sse2_ms.append(pr.timer_stop())
        ps = pr.profile()
        pr.profile_enable(False)
        sse2 = {"ms_per_step": sum(sse2_ms) / len(sse2_ms), "planes": len(exs.planes),
                "score_tflops": 6.0 * ps.pairs_scored / (ps.ms_score * 1e-3) / 1e12 if ps.ms_score > 0 else None,
                "note": "PR_DOT_PCL_SSE2: (a*x + c*z) + (b*y + d) with separately rounded products, the order PCL 1.8 / MSVC v140 "
                        "evaluates; 6 FP32 operations per point-hypothesis issue as 6 scalar instructions (ptxas fuses packed mul+add "
                        "into FFMA2, which would change the bits), against 1.5 FFMA2 for the FMA order; this rank's figure"}

    # ---- same workload through the opt-in hierarchical (bounding-box culled) scorer: identical planes ----
    hier_ms = []
    hier_same = None
    if args.scorer == "brute":
        prm_h = D.make_params(0.1, args.hyps - 1, 500, 1.0, True, 12345, args.planes, D.DOT_FMA, D.SCORER_HIER)
        for _ in range(2):
            exh = pr.extract_planes(prm_h, want_indices=False)
        barrier()
        for _ in range(args.steps):
            pr.flush_l2()
            barrier()
            pr.timer_start()
            exh = pr.extract_planes(prm_h, want_indices=False)
            hier_ms.append(pr.timer_stop())
        barrier()
        hier_same = len(exh.planes) == len(ex.planes) and all(
            a.coeff.tobytes() == b.coeff.tobytes() and a.info.n_inliers == b.info.n_inliers
            for a, b in zip(exh.planes, ex.planes))
    hier_total_ms = sum(hier_ms)

    # ---- per-round, per-rank timeline of the last timed step (the ranks' clocks are not synchronised with each other: each
    #      rank's stamps are relative to its own first one) ----
    if args.timeline_out:
        rel = [[int(v) - int(timeline[0][0]) if v else None for v in row] for row in timeline.tolist()] if len(timeline) else []
        per_rank = [rel]
        if world > 1:
            per_rank = [None] * world
            dist.all_gather_object(per_rank, rel)
        if rank == 0:
            with open(args.timeline_out, "w") as f:
                json.dump({"n_gpus": world, "points_per_gpu": count, "hypotheses_per_round": args.hyps,
                           "stages": list(D.LOOP_STAGE_NAMES) + ["record_written"],
                           "unit": "ns since the rank's own first stamp (the GPUs' %globaltimer values are not aligned with each "
                                   "other); null: the round has no such stage",
                           "ranks": per_rank}, f)

    # ---- per-rank time spent inside the exchange kernels waiting for the peers (lag of the slowest rank + NVLink round trip) ----
    waits = None
    if world > 1:
        w = torch.tensor([list(prof.p2p_wait_ms) + [float(x) for x in prof.p2p_exchanges]], dtype=torch.float64, device="cuda")
        allw = [torch.zeros_like(w) for _ in range(world)]
        dist.all_gather(allw, w)
        waits = [[v / args.steps for v in x.flatten().tolist()] for x in allw]

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_total_ms, hier_total_ms, db_total_ms, statistics.median(up_ms)], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms, e2e_total_ms, hier_total_ms, db_total_ms, up_alone_ms = t.tolist()
    h2d_all, d2h_all = agg.tolist()

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = pairs_step / (ms_per_step * 1e-3)
        e2e_value = pairs_step / (e2e_total_ms / args.steps * 1e-3)
        peaks = measured_peaks()
        hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
        # scoring time inside the timed region: device stamps (start of K2 to start of the kernel after it) when the
        # rounds ran in the device-resident loop, else the event pass
        loop_ms = list(prof_t.loop_ms)
        stamped = prof_t.loop_rounds > 0 and loop_ms[3] > 0
        score_ms_timed = loop_ms[3] if stamped else prof.ms_score
        score_pairs = prof_t.pairs_scored if stamped else prof.pairs_scored
        score_tf = 6.0 * score_pairs / (score_ms_timed * 1e-3) / 1e12 if score_ms_timed > 0 else None
        score_tf_events = 6.0 * prof.pairs_scored / (prof.ms_score * 1e-3) / 1e12 if prof.ms_score > 0 else None
        launches = (prof_t.launches_stage + prof_t.launches_models + prof_t.launches_score + prof_t.launches_refit +
                    prof_t.launches_compact + prof_t.launches_other)
        kernel_ms = prof.ms_models + prof.ms_score + prof.ms_refit + prof.ms_compact
        ev_ms_per_step = sum(ev_ms) / len(ev_ms)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world, pr.p2p_enabled()),
            "planes_extracted": n_planes, "pairs_per_step": pairs_step,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_total_ms / args.steps,
                    "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all)},
            "e2e_breakdown": {
                "upload_alone_ms": up_alone_ms, "upload_gbs_per_rank": count * 16 / (up_alone_ms * 1e-3) / 1e9,
                "upload_note": "plane_ransac_set_cloud (blocking copy + staging kernel) of every rank's %d MB at the same time, slowest rank; "
                               "in e2e the copy is chunked and hidden under round 0's scoring as far as that lasts" % (count * 16 // 2**20),
                "e2e_minus_resident_ms": e2e_total_ms / args.steps - ms_per_step,
                "double_buffered": {
                    "ms_per_step": db_total_ms / args.steps, "value": pairs_step / (db_total_ms / args.steps * 1e-3), "unit": UNIT,
                    "note": "the same end-to-end steps (pinned cloud in, coefficients + index lists out, every step) alternating two "
                            "contexts per GPU: step s + 1's upload is queued before step s's extraction starts and travels under it; "
                            "host wall clock between barriers; no L2 flush inside (each step's 160 MB arrive by DMA meanwhile)"}},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": "score_kernel<8,FMA>" if args.scorer == "brute" else
                          "score_hier_kernel (culled: 'achieved' counts the point-hypotheses decided, not the FMAs executed, "
                          "so frac is not a pipe utilisation)",
                "bound": "fp32_fma", "achieved": score_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": (score_tf / peak_tf) if score_tf and peak_tf else None,
                "traffic": (ncu_traffic() or {}).get("score_kernel_dram_bytes_per_launch"),
                "traffic_note": (ncu_traffic() or {}).get("note"),
                "peak_source": "FFMA2-only kernel timed live on this GPU (MEASURED_PEAKS.json has no FP32 figure; "
                               "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)",
                "algorithmic": "6 FLOP (3 FMA) per point-hypothesis x points x hypotheses per launch",
                "ms_in_timed_region": score_ms_timed,
                "timing": ("%globaltimer stamps taken by the kernels themselves inside the timed region: start of the scoring "
                           "launch to start of the kernel that follows it (the kernel plus one hand-over), summed over the "
                           "rounds of the timed steps") if stamped else "CUDA events around the scoring launches (event pass)",
                "share_of_step": score_ms_timed / total_ms if total_ms else None,
                "frac_cuda_events": (score_tf_events / peak_tf) if score_tf_events and peak_tf else None,
                "frac_cuda_events_note": "the same launches timed with CUDA events on the launching stream in the event pass "
                                         "(same steps run again outside the timed region)",
                "share_of_kernel_time": prof.ms_score / kernel_ms if kernel_ms else None},
            "roofline_hbm": dict(hbm_block(prof, hbm_peak),
                                 peak_source="MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                                 note="averages over the 20 shrinking rounds of the 10M-point extraction (the later rounds are "
                                      "launch-latency bound); single_launch_100M below is the kernel at the north-star scene size"),
            "device_loop_ms_per_step": dict(
                {nm: loop_ms[i] / args.steps for i, nm in enumerate(D.LOOP_STAGE_NAMES)},
                rounds_per_step=prof_t.loop_rounds / args.steps,
                rounds_total=sum(loop_ms) / args.steps,
                step_minus_rounds=ms_per_step - sum(loop_ms) / args.steps,
                note="timed region, this rank: %globaltimer at the start of each kernel of a round (after its predecessor "
                     "completed) and when the round's record was written; a stage is its kernel plus the hand-over to the "
                     "next one, so the stages add up to the rounds' device time and step_minus_rounds is what the host "
                     "adds (first launch, last record, the final sync)") if stamped else None,
            "kernel_ms_per_step": {"models_draw": prof.ms_models / args.steps, "score": prof.ms_score / args.steps,
                                   "refit": prof.ms_refit / args.steps, "compact": prof.ms_compact / args.steps,
                                   "other": prof.ms_other / args.steps,
                                   "event_pass_ms_per_step": ev_ms_per_step,
                                   "step_minus_kernels": ev_ms_per_step - (kernel_ms + prof.ms_other) / args.steps,
                                   "note": "event pass (outside the timed region): CUDA events around every kernel class; "
                                           "step_minus_kernels is against that pass's own step time"},
            "round_loop": "device-resident (PR_LOOP_AUTO): draws, computeModel's decision, closed-form refit and the stop rule run "
                          "as kernels; the host reads one record per round",
            "clocks": clock_info,
        }
        if sse2:
            sse2["frac_of_fp32_peak"] = sse2["score_tflops"] / peak_tf if sse2["score_tflops"] else None
            line["pcl_sse2_dot_order"] = sse2
        if waits:
            line["exchange_wait_ms_per_step_by_rank"] = {
                "channels": ["sample points", "counts", "refit moments", "remaining counts"],
                "wait_ms": [x[:4] for x in waits], "exchanges_per_step": waits[0][4:],
                "note": "time each rank's exchange kernels spin on the peers' flags inside the resident timed region: a rank that "
                        "finishes its kernels early waits for the slowest one (straggling), every rank pays the NVLink round trip"}
        if hier_ms:
            line["hier_scorer"] = {
                "ms_per_step": hier_total_ms / args.steps, "value": pairs_step / (hier_total_ms / args.steps * 1e-3),
                "unit": UNIT, "identical_planes": bool(hier_same),
                "note": "opt-in PR_SCORER_HIER: Morton-sorted copy + per-32-point boxes, blocks outside the threshold slab "
                        "skipped, the rest evaluated with the same arithmetic; same counts, not the roofline kernel"}
        if args.extras:
            line["extras"] = run_extras(args, pr)
    # ---- BASELINE configs[3]: the 100M-point scene sharded over the N ranks, outside the headline timer ----
    pr.close()
    c4 = None
    if not args.no_config4:
        c4 = run_config4(args, torch, dist, D, local_rank, world, rank, hbm_peak if rank == 0 else None, peak_tf)
    c5 = None
    if not args.no_config5:
        c5 = run_config5(args, torch, dist, D, local_rank, world, rank, peak_tf)
    if rank == 0:
        line["config2_1M"] = run_config2(D, local_rank, peak_tf) if not args.no_config5 else None
        line["config5_batch"] = c5
        if c4 and "single_launch" in c4:
            # the HBM kernels at the north star's scene size, one launch each, timed alone; the figures in roofline_hbm
            # proper are averages over the shrinking 10M-point rounds
            line["roofline_hbm"]["single_launch_%dM_points" % (c4["points_total"] // 1_000_000)] = c4.pop("single_launch")
        line["config4_100M"] = c4
        if not args.no_cpu_baseline:
            v, dt = cpu_sample(args, pts, threads=1)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
                "sample": "oracle countWithinDistance (PCL 1.8 scalar loop, 1 thread) over the first %d hypotheses of "
                          "round 0 on rank 0's %d points" % (args.cpu_sample_hyps, count)}
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
