"""Host -> device bandwidth per rank when N ranks copy at once (torchrun): copy-engine transfers of page-locked memory
(cudaMemcpyAsync) next to the zero-copy fetch of fuz_phase_batch_host.   torchrun --nproc-per-node N scripts/h2d_bw_ranks.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    res = {}
    for name, chunks in (("dma_1GiB", 1), ("dma_64MiB_chunks", 16)):
        for _ in range(2):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            step = n // chunks
            for c in range(chunks):
                d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        res[name] = reps * n / (time.perf_counter() - t0) / 1e9
    out = torch.tensor([res["dma_1GiB"], res["dma_64MiB_chunks"]], device="cuda", dtype=torch.float64)
    allv = [torch.zeros_like(out) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, out)
    else:
        allv = [out]
    if rank == 0:
        print(json.dumps({"n_ranks": world, "dma_GBps_per_rank": [round(float(v[0]), 1) for v in allv],
                          "dma_chunked_GBps_per_rank": [round(float(v[1]), 1) for v in allv],
                          "sum_GBps": round(sum(float(v[0]) for v in allv), 1)}))
    if world > 1:
        dist.destroy_process_group()


main()
