"""Platform ceiling for the end-to-end number: every rank copies a 2.07 GB device buffer (1000 index frames at 1080p) to
pinned host memory, all ranks at once -- no rendering, no library.  What this prints per rank is the most `e2e` can be on
this box at that GPU count.  Run under torchrun like bench.py; `--no-numa` leaves the process unbound."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--no-numa", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = None
    if not args.no_numa:
        from rust_doom_b200 import jobs
        numa = jobs.bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("gloo")
    nbytes = 1000 * 1920 * 1080
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for h in host:
        h.fill_(1)                                             # first touch on this rank's node
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    def one_pass(reps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        start.record()
        for s in streams:
            s.wait_event(start)
        for k in range(reps):
            with torch.cuda.stream(streams[k & 1]):
                host[k & 1].copy_(dev, non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end)
    one_pass(2)
    ms = one_pass(args.reps)
    gbs = args.reps * nbytes / ms / 1e6
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, {"rank": rank, "gbs": gbs, "numa": numa})
    else:
        out = [{"rank": rank, "gbs": gbs, "numa": numa}]
    if rank == 0:
        g = [o["gbs"] for o in out]
        print(json.dumps({"what": "D2H ceiling, all ranks at once, 2.07 GB per copy, two streams per rank", "n_gpus": world, "numa_bound": not args.no_numa,
                          "gbs_per_rank": [round(x, 2) for x in g], "min": round(min(g), 2), "sum": round(sum(g), 1),
                          "frames_per_s_ceiling": round(min(g) * 1e9 / (1920 * 1080) * world, 0), "numa": [o["numa"] for o in out]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
