"""scratch: N-GPU step time with / without the in-library all-reduce (torchrun)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import circuitsimulator_b200 as bg
import bench
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, Gd, Hd, samples, k, desc = bench.load_config("hidden_shift_n40_t40_k9_L65536")
t = cfg["t"]; L = bench.fixed_L(k, t)
G = bg.Projector.make(*Gd); H = bg.Projector.make(*Hd)
for mode in ("allreduce", "no_allreduce"):
    ctx = bg.Backend(local)
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    uid = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.set_stream(ts.cuda_stream); ctx.set_shard(rank, world); ctx.nccl_join(uid[0])
    if mode == "no_allreduce":
        ctx._ck(ctx.lib.bg_set_allreduce(ctx.ctx, 0))
    ctx.set_decomposition(t, 0, L)
    ctx.sampled_prepare2(G, H, samples, 1, 1001, 1002)
    for _ in range(5):
        ctx.sampled_run(); ctx.sampled_finish2(1.0)
    dist.barrier(); torch.cuda.synchronize()
    steps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    ctx.sampled_run()
    km = []
    for _ in range(steps - 1):
        ctx.sampled_run(); ctx.sampled_finish2(1.0); km.append(ctx.stats()["kernel_ms"])
    ctx.sampled_finish2(1.0)
    e1.record(ts)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    allms = [None] * world
    dist.all_gather_object(allms, (ms, sorted(km)[len(km) // 2]))
    if rank == 0:
        print(mode, "per-rank ms/step and median graph ms:", json.dumps([[round(a, 4), round(b, 4)] for a, b in allms]), flush=True)
    ctx.close()
    dist.barrier()
dist.destroy_process_group()
