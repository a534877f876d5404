"""CPU, world_size 2, gloo: the host-side protocol of the point-sharded path.

Each rank owns a contiguous shard (plane_ransac_host_shard_range).  The oracle stands in for the device
kernels (it is the checker, never the product): per-shard inlier counts are summed with an all-reduce,
the sample points are exchanged as integer bit patterns (owner contributes, others zero), the library's
replay picks the winner, per-shard integer moments are all-reduced and the library turns them into the
plane.  Two peel rounds are run so that the per-rank prefix of the peeled cloud (global index = prefix + local index) is
exercised.  Everything must equal the single-process oracle on the whole cloud, bit for bit."""
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
        shard = pts[first: first + count]            # this rank's current cloud (peeled round by round)
        K, t, max_it, rounds = 200, 0.1, 199, 2
        s = O.fixed_scale_exp(pts)                   # global bounding box (the device path all-reduces min/max keys)
        want = O.extract_planes(pts, O.make_params(t, max_it, 500, 1.0, True, 12345, rounds, O.DOT_FMA, O.REFIT_FIXED))
        assert len(want.coeffs) == rounds
        orig = np.arange(first, first + count)       # original index of every point still in the shard
        n_cur = n
        for rnd in range(rounds):
            # global size and this rank's first global index in the CURRENT (peeled) cloud: all-gather of shard sizes
            sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([shard.shape[0]], dtype=torch.int64))
            sizes = [int(v.item()) for v in sizes]
            n_cur, first_cur = sum(sizes), sum(sizes[:rank])
            # 1. identical draw stream on every rank (indices into the current global cloud)
            tri = D.host_draw_triples(n_cur, K)
            # 2. sample points: owner contributes the bits, others zero; integer sum == exact transfer
            flat = tri.reshape(-1)
            mine = (flat >= first_cur) & (flat < first_cur + shard.shape[0])
            bits = np.zeros((flat.size, 3), np.int32)
            bits[mine] = shard[flat[mine] - first_cur, :3].view(np.int32)
            tb = torch.from_numpy(bits)
            dist.all_reduce(tb)
            sample_pts = tb.numpy().view(np.float32)
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
            rep = D.host_replay(counts, good, n_cur, max_it, 1.0)
            best = rep["best_draw"]
            # 5. refit: per-shard exact integer moments about the winner's first sample point, summed
            pivot = sample_pts[3 * best]
            idx = O.select_within(shard, coeffs[best], t, O.DOT_FMA)
            _, mom = O.refit_fixed(shard, idx, pivot, s, coeffs[best])
            if idx.size < 4:           # refit_fixed zeroes the moments of tiny shards; accumulate them anyway
                mom = np.zeros(16, np.int64)
            tm = torch.from_numpy(mom.copy())
            dist.all_reduce(tm)
            refined = D.host_plane_from_moments(tm.numpy(), pivot, s)
            # 6. final selection on the shard; local peel, order preserved
            inl = O.select_within(shard, refined, t, O.DOT_FMA)
            tn = torch.tensor([inl.size])
            dist.all_reduce(tn)
            # compare with round `rnd` of the single-process oracle on the whole cloud
            assert tri[best].tolist() == list(want.traces[rnd].best_sample)
            assert counts[best] == want.traces[rnd].best_count and rep["iterations"] == want.traces[rnd].iterations
            assert refined.tobytes() == want.coeffs[rnd].tobytes()
            assert int(tn.item()) == want.inliers_orig[rnd].size
            w = want.inliers_orig[rnd]
            assert (orig[inl] == w[(w >= first) & (w < first + count)]).all()
            # local-current index + this rank's prefix == PCL's index into the current cloud
            assert np.isin(inl + first_cur, want.inliers_cur[rnd]).all()
            keep = np.ones(shard.shape[0], bool)
            keep[inl] = False
            shard, orig = shard[keep], orig[keep]
        rest = want.remaining
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([shard.shape[0]], dtype=torch.int64))
        assert sum(int(v.item()) for v in sizes) == rest.shape[0]
        # postProcessPlanes re-absorption, sharded: the predicate is per point, so every rank claims on its own shard
        # with no data exchange; only the remaining counts are gathered (the new prefix of the peeled cloud)
        scene = synth.three_planes_scene()
        borders = []
        for c in want.coeffs:
            err = [min(np.abs(q.coeff - c).max(), np.abs(q.coeff + c).max()) for q in scene.patches]
            borders.append(scene.patches[int(np.argmin(err))].border(10))
        whole = O.reabsorb(rest, want.coeffs, borders, 0.2, 31)
        mine = O.reabsorb(shard, want.coeffs, borders, 0.2, 31)
        sizes = [int(v.item()) for v in sizes]
        first_cur = sum(sizes[:rank])
        for k in range(rounds):
            w = whole.absorbed[k]
            assert np.array_equal(mine.absorbed[k] + first_cur, w[(w >= first_cur) & (w < first_cur + shard.shape[0])])
        left = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(left, torch.tensor([mine.remaining_idx.size], dtype=torch.int64))
        assert sum(int(v.item()) for v in left) == whole.remaining_idx.size
        assert sum(len(a) for a in whole.absorbed) > 20
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
