"""N-GPU check + timing of mdg_peer_allgather against NCCL all_gather_into_tensor (launch with torchrun)."""
import os, sys, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from madrigal_b200 import scoring

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for N, D in ((4096, 256), (4099, 256), (20000, 256)):
    g = scoring.PeerAllGather(N, D, dev)
    r0, r1 = scoring.row_shard(N, rank, world)
    ok = True
    for it in range(6):
        torch.manual_seed(100 * it + rank)
        shard = torch.randn(r1 - r0, D, device=dev)
        got = g.gather(shard).clone()
        ref = scoring.all_gather_embeddings(shard, N)
        ok &= bool(torch.equal(got, ref))
    def timed(fn, iters=50):
        for _ in range(5): fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3
    shard = torch.randn(r1 - r0, D, device=dev)
    out = torch.empty(N, D, device=dev) if N % world == 0 else None
    us_peer = timed(lambda: g.gather(shard))
    us_nccl = timed(lambda: scoring.all_gather_embeddings(shard, N, out=out))
    res[f"{N}x{D}"] = {"mode": g.mode, "reason": g.reason, "equal_to_nccl": ok, "peer_us": us_peer, "nccl_us": us_nccl}
flag = torch.tensor([1 if all(v["equal_to_nccl"] for v in res.values()) else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    res["all_ranks_equal"] = bool(flag.item()); res["world"] = world
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"peer_allgather_{world}gpu.json"), "w"), indent=1)
dist.destroy_process_group()
