"""Sharded (torchrun, N ranks): where the end-to-end overhead of an extraction goes — resident against uploaded cloud,
with and without index lists, 1 / 2 / 20 planes."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
import dialog_b200 as D
from dialog_b200 import synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 10_000_000
first, count = D.host_shard_range(n * world, world, rank)
pin = D.PinnedArray((count, 4), np.float32)
pin.array[:] = synth.indoor_scene().points(first, first + count)
pr = D.PlaneRansac(local)
uid = [D.PlaneRansac.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
pr.comm_init(world, rank, uid[0])

def barrier():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

for planes in (1, 2, 20):
    prm = D.make_params(0.1, 4095, 500, 1.0, True, 12345, planes, D.DOT_FMA)
    for mode in ("resident", "uploaded", "uploaded+lists", "sync-upload+lists"):
        ms, host = [], []
        for rep in range(6):
            if mode == "resident":
                pr.set_cloud_ptr(pin.ptr, count)
            pr.flush_l2()
            barrier()
            pr.profile_reset()
            t0 = time.perf_counter()
            pr.timer_start()
            if mode.startswith("uploaded"):
                pr.set_cloud_ptr(pin.ptr, count, overlap=True)
            elif mode.startswith("sync"):
                pr.set_cloud_ptr(pin.ptr, count)
            t1 = time.perf_counter()
            ex = pr.extract_planes(prm, want_indices=mode.endswith("lists"), copy=False)
            t2 = time.perf_counter()
            ms.append(pr.timer_stop())
            host.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
            p = pr.profile()
        t = torch.tensor([float(np.median(ms[2:]))], device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        if rank == 0:
            print(f"planes={planes} {mode}: per rank ms {[round(x.item(), 3) for x in allt]}; rank 0 host: set_cloud {host[-1][0]:.3f} extract {host[-1][1]:.3f} "
                  f"sampling {p.host_ms_sampling:.3f} replay {p.host_ms_replay:.3f} wait {p.host_ms_wait:.3f} total {p.host_ms_total:.3f} loop rounds {p.loop_rounds}", flush=True)
pr.close()
dist.destroy_process_group()
