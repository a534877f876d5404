"""Launched under torchrun by tests/test_gpu_multi.py (and by hand on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank stages a contiguous shard of the same cloud; the sharded extraction (NCCL all-reduce of counts and
moments) must reproduce the single-GPU extraction of the whole cloud bit for bit: same coefficients, same
per-plane inlier sets (local indices + shard offset == global indices), same remaining points."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dialog_b200 as D  # noqa: E402
from dialog_b200 import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("PR_CHECK_POINTS", "600000"))
    pts = synth.indoor_scene().points(0, n)
    prm = D.make_params(0.1, 1023, 500, 1.0, True, 12345, 20, D.DOT_FMA)

    # single-GPU answer (every rank computes it on its own device)
    with D.PlaneRansac(local) as one:
        one.set_cloud(pts)
        want = one.extract_planes(prm)
        want_rem = one.remaining().copy()
        want_counts = one.score(D.host_draw_triples(n, 300), 0.1)

    first, count = D.host_shard_range(n, world, rank)
    sh = D.PlaneRansac(local)
    uid = [D.PlaneRansac.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sh.comm_init(world, rank, uid[0])
    want_p2p = os.environ.get("PR_P2P", "1") != "0"
    assert sh.p2p_enabled() == want_p2p, f"peer-memory exchanges: enabled={sh.p2p_enabled()}, expected {want_p2p}"
    sh.set_cloud(pts[first: first + count])
    assert sh.shard_info()[:2] == (n, first)
    got_counts = sh.score(D.host_draw_triples(n, 300), 0.1)
    assert (got_counts == want_counts).all(), "sharded counts differ"
    got = sh.extract_planes(prm)
    assert len(got.planes) == len(want.planes), (len(got.planes), len(want.planes))
    for k, (a, b) in enumerate(zip(got.planes, want.planes)):
        assert a.coeff.tobytes() == b.coeff.tobytes(), f"plane {k}: coefficients differ"
        assert a.info.n_inliers == b.info.n_inliers and list(a.info.best_sample) == list(b.info.best_sample)
        mine = b.inliers_orig[(b.inliers_orig >= first) & (b.inliers_orig < first + count)] - first
        assert (a.inliers_orig == mine).all(), f"plane {k}: inlier set differs on rank {rank}"
    # the host-driven loop (three synchronisations per round) gives the same planes as the device-driven default
    sh.set_round_loop(host=True)
    got_hl = sh.extract_planes(prm)
    sh.set_round_loop(host=False)
    assert len(got_hl.planes) == len(got.planes)
    for k, (a, b) in enumerate(zip(got_hl.planes, got.planes)):
        assert a.coeff.tobytes() == b.coeff.tobytes() and (a.inliers_orig == b.inliers_orig).all(), f"host loop, plane {k}"
        assert list(a.info.best_sample) == list(b.info.best_sample) and a.info.best_count == b.info.best_count
    # the hierarchical scorer, sharded: every rank culls on its own Morton-sorted shard, same global answer
    got_h = sh.extract_planes(D.make_params(0.1, 1023, 500, 1.0, True, 12345, 20, D.DOT_FMA, D.SCORER_HIER))
    assert len(got_h.planes) == len(want.planes)
    for k, (a, b) in enumerate(zip(got_h.planes, got.planes)):
        assert a.coeff.tobytes() == b.coeff.tobytes() and (a.inliers_orig == b.inliers_orig).all(), f"hier plane {k}"
    rem = sh.remaining()
    claimed = np.zeros(n, bool)
    claimed[np.concatenate([p.inliers_orig for p in want.planes])] = True
    assert rem.tobytes() == pts[first: first + count][~claimed[first: first + count]].tobytes()
    tot = torch.tensor([rem.shape[0]], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == want_rem.shape[0]
    # unbalanced shards: rank 0 holds one whole 300k-point plane plus a sliver of the scene, so after the first peel its
    # shard is a fraction of the others'.  Everything a rank derives its collectives from (sub-batch sizes, exchange
    # mode) must come from global quantities; 8192 hypotheses per round make the host-driven loop cut sub-batches.
    rng = np.random.default_rng(5)
    slab = np.ones((300_000, 4), np.float32)
    slab[:, 0] = rng.uniform(0, 30, 300_000)
    slab[:, 1] = rng.uniform(0, 20, 300_000)
    slab[:, 2] = 9.0 + rng.normal(0, 0.02, 300_000)
    lop = np.concatenate([slab, pts[:300_000]])
    cuts = [0, 330_000] + [330_000 + (270_000 * (r + 1)) // (world - 1) for r in range(world - 1)]
    prm_u = D.make_params(0.1, 8191, 500, 1.0, True, 12345, 6, D.DOT_FMA)
    with D.PlaneRansac(local) as one:
        one.set_cloud(lop)
        want_u = one.extract_planes(prm_u)
    assert want_u.planes[0].info.n_inliers >= 300_000
    dist.barrier()
    f_u, c_u = cuts[rank], cuts[rank + 1] - cuts[rank]
    sh.set_cloud(lop[f_u: f_u + c_u])
    assert sh.shard_info()[:2] == (lop.shape[0], f_u)
    for host_loop in (False, True):
        sh.set_round_loop(host=host_loop)
        got_u = sh.extract_planes(prm_u)
        assert len(got_u.planes) == len(want_u.planes) == 6
        for k, (a, b) in enumerate(zip(got_u.planes, want_u.planes)):
            assert a.coeff.tobytes() == b.coeff.tobytes() and a.info.n_inliers == b.info.n_inliers, f"unbalanced, plane {k}"
            mine = b.inliers_orig[(b.inliers_orig >= f_u) & (b.inliers_orig < f_u + c_u)] - f_u
            assert (a.inliers_orig == mine).all(), f"unbalanced shards, plane {k}: inlier set differs on rank {rank}"
    sh.set_round_loop(host=False)
    # overlapped upload, sharded: chunks scored as they land on every rank, same answer (shards of >= 2M points)
    n_big = 2_200_000 * world
    big = synth.indoor_scene().points(0, n_big)
    prm_b = D.make_params(0.1, 1023, 500, 1.0, True, 12345, 3, D.DOT_FMA)
    with D.PlaneRansac(local) as one:
        one.set_cloud(big)
        want_b = one.extract_planes(prm_b)
    fb, cb = D.host_shard_range(n_big, world, rank)
    pin = D.PinnedArray((cb, 4), np.float32)
    pin.array[:] = big[fb: fb + cb]
    dist.barrier()  # the ranks generated and solved the big cloud on their own: line them up before the collective calls
    sh.set_cloud_ptr(pin.ptr, cb, overlap=True)
    assert sh.shard_info()[:2] == (n_big, fb)
    got_b = sh.extract_planes(prm_b)
    assert len(got_b.planes) == len(want_b.planes) == 3
    for k, (a, b) in enumerate(zip(got_b.planes, want_b.planes)):
        assert a.coeff.tobytes() == b.coeff.tobytes() and a.info.scale_exp == b.info.scale_exp, f"async upload, plane {k}"
        mine = b.inliers_orig[(b.inliers_orig >= fb) & (b.inliers_orig < fb + cb)] - fb
        assert (a.inliers_orig == mine).all(), f"async upload, plane {k}: inlier set differs on rank {rank}"
    pin.free()
    # re-absorption pass, sharded: every rank claims on its shard; the union is the single-GPU answer
    scene = synth.indoor_scene()
    prm_t = D.make_params(0.05, 200, 500, 0.99, True, 12345, 6, D.DOT_FMA)
    with D.PlaneRansac(local) as one:
        one.set_cloud(pts)
        ex_t = one.extract_planes(prm_t)
        coeffs = np.array([p.coeff for p in ex_t.planes], np.float32)
        borders = []
        for c in coeffs:
            err = [min(np.abs(q.coeff - c).max(), np.abs(q.coeff + c).max()) for q in scene.patches]
            borders.append(scene.patches[int(np.argmin(err))].border(12))
        _, want_orig, want_left = one.reabsorb(coeffs, borders, 0.1, 9)
    dist.barrier()
    sh.set_cloud(pts[first: first + count])
    got_t = sh.extract_planes(prm_t)
    assert [p.coeff.tobytes() for p in got_t.planes] == [p.coeff.tobytes() for p in ex_t.planes]
    _, got_orig, got_left = sh.reabsorb(coeffs, borders, 0.1, 9)
    for k in range(len(coeffs)):
        mine = want_orig[k][(want_orig[k] >= first) & (want_orig[k] < first + count)] - first
        assert (got_orig[k] == mine).all(), f"re-absorption, plane {k}, rank {rank}"
    tot = torch.tensor([got_left], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == want_left and sh.shard_info()[2] == want_left
    assert sum(len(a) for a in want_orig) > 50
    # fused preProcess staging, sharded: the centroid is the global one (integer sums all-reduced), so the
    # translated shards and the planes found on them equal the single-GPU preprocessed run
    dirty = pts.copy()
    dirty[::1013, 1] = np.nan
    dirty[:, :3] += np.float32(100.0)
    with D.PlaneRansac(local) as one:
        kept1, cen1, src1 = one.set_cloud_preprocessed(dirty)
        want_p = one.extract_planes(prm)
    kept, cen, src = sh.set_cloud_preprocessed(dirty[first: first + count])
    assert cen.tobytes() == cen1.tobytes(), (cen, cen1)
    tot = torch.tensor([kept], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == kept1
    got_p = sh.extract_planes(prm)
    assert len(got_p.planes) == len(want_p.planes)
    for a, b in zip(got_p.planes, want_p.planes):
        assert a.coeff.tobytes() == b.coeff.tobytes() and a.info.n_inliers == b.info.n_inliers
    # BASELINE configs[4]: a batch of 32K-point tiles sharded by cloud id (cloud_id % world == rank): replicas only, no
    # collective — every rank checks its 64 clouds against the per-cloud answer of a plain single-GPU context
    from oracle import oracle as O
    ids = [cid for cid in range(64 * world) if cid % world == rank]
    tiles = np.stack([synth.tile_scene(cid).points(0, 32768) for cid in ids])
    prm_b5 = D.make_params(0.1, 255, 500, 1.0, True, 12345, 1, D.DOT_FMA)
    with D.PlaneRansac(local) as tb:
        tb.set_cloud_batch(tiles)
        bc, bn, bi, bl = tb.segment_batch(prm_b5, want_lists=True)
    for j, cid in enumerate(ids[:: max(1, len(ids) // 16)]):     # 16 of the 64 against the CPU oracle (the rest: see tests/test_gpu_parity.py)
        k = ids.index(cid)
        seg = O.segment(tiles[k], O.make_params(0.1, 255, 500, 1.0, True, 12345, 1, O.DOT_FMA, O.REFIT_FIXED))
        assert seg.ok and bc[k].tobytes() == seg.coeff.tobytes() and np.array_equal(bl[k], seg.inliers), f"batch cloud {cid} on rank {rank}"
    sh.close()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check ok: {world} ranks, {len(got.planes)} planes, bit-identical to one GPU, "
              f"exchanges: {'peer-memory kernels' if want_p2p else 'NCCL'}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
