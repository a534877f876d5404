"""CPU, world_size 2, gloo: the host-side protocol of the point-sharded path.

Each rank owns a contiguous shard (plane_ransac_host_shard_range).  The oracle stands in for the device
kernels (it is the checker, never the product): per-shard inlier counts are summed with an all-reduce,
the sample points are exchanged as integer bit patterns (owner contributes, others zero), the library's
replay picks the winner, per-shard integer moments are all-reduced and the library turns them into the
plane.  Everything must equal the single-process oracle on the whole cloud, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dialog_b200 as D
        from dialog_b200 import synth
        from oracle import oracle as O

        pts = synth.three_planes_scene().points(0, n)
        first, count = D.host_shard_range(n, world, rank)
        shard = pts[first: first + count]
        K, t, max_it = 200, 0.1, 199

        # 1. identical draw stream on every rank (global indices)
        tri = D.host_draw_triples(n, K)
        # 2. sample points: owner contributes the bits, others zero; integer sum == exact transfer
        flat = tri.reshape(-1)
        mine = (flat >= first) & (flat < first + count)
        bits = np.zeros((flat.size, 3), np.int32)
        bits[mine] = shard[flat[mine] - first, :3].view(np.int32)
        tb = torch.from_numpy(bits)
        dist.all_reduce(tb)
        sample_pts = tb.numpy().view(np.float32)
        assert sample_pts.tobytes() == pts[flat, :3].tobytes()
        # 3. models from the gathered points (same on every rank), counts on the shard, all-reduce
        cloud9 = np.ones((flat.size, 4), np.float32)
        cloud9[:, :3] = sample_pts
        coeffs, good = O.models_from_triples(cloud9, np.arange(flat.size, dtype=np.int32).reshape(-1, 3))
        local = O.count_batch(shard, np.nan_to_num(coeffs), t, O.DOT_FMA)
        tc = torch.from_numpy(local.copy())
        dist.all_reduce(tc)
        counts = tc.numpy()
        counts[~good] = 0
        # 4. the library replays PCL's sequential decisions over the summed counts
        rep = D.host_replay(counts, good, n, max_it, 1.0)
        best = rep["best_draw"]
        # 5. refit: per-shard exact integer moments about the winner's first sample point, summed
        s = O.fixed_scale_exp(pts)     # global bounding box (the device path all-reduces min/max keys)
        pivot = sample_pts[3 * best]
        idx = O.select_within(shard, coeffs[best], t, O.DOT_FMA)
        _, mom = O.refit_fixed(shard, idx, pivot, s, coeffs[best])
        if idx.size < 4:               # refit_fixed zeroes the moments of tiny shards; accumulate them anyway
            mom = np.zeros(16, np.int64)
        tm = torch.from_numpy(mom.copy())
        dist.all_reduce(tm)
        refined = D.host_plane_from_moments(tm.numpy(), pivot, s)
        # 6. final selection on the shard; global inlier count by all-reduce
        inl = O.select_within(shard, refined, t, O.DOT_FMA)
        tn = torch.tensor([inl.size])
        dist.all_reduce(tn)

        # single-process oracle on the whole cloud
        seg = O.segment(pts, O.make_params(t, max_it, 500, 1.0, True, 12345, 8, O.DOT_FMA, O.REFIT_FIXED))
        assert tri[best].tolist() == list(seg.trace.best_sample)
        assert counts[best] == seg.trace.best_count and rep["iterations"] == seg.trace.iterations
        assert refined.tobytes() == seg.coeff.tobytes()
        assert int(tn.item()) == seg.inliers.size
        want_local = seg.inliers[(seg.inliers >= first) & (seg.inliers < first + count)] - first
        assert (inl == want_local).all()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [20_001])
def test_sharded_protocol_matches_single_process(lib_built, O, n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
