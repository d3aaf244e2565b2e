"""Per-kernel device time of one training step of a bench workload via torch.profiler (CUPTI), low overhead.
Usage: python tools/profile_step.py [workload]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    import os
    import torch.distributed as dist
    from dexbotic_b200.parallel import GradientOverlap, ShardedDataParallel
    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cogact_7b"]
    local, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                      # torchrun: profile rank 0 of a data-parallel step (NCCL kernels included)
        dist.init_process_group("nccl", device_id=dev)
    model = bench.build_model(w, dev)
    overlap = (ShardedDataParallel(model.store) if world > 1 and os.environ.get("B200_DP", "zero1") == "zero1"
               else GradientOverlap(model.store))
    model.store.async_optimizer = True
    model.init_weights_(seed=1234)
    model.train()
    batch = {k: (v.to(dev) if hasattr(v, "to") else v) for k, v in bench.make_batch(w, 0, pinned=False).items()}

    def step():
        model.zero_grad()
        out = model(**batch)
        out.loss.backward()
        overlap.finish()
        model.optimizer_step(base_lr=2e-5)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
            rows.append((t, e.count, e.key))
    if local != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"total device time {tot/1e3:.1f} ms over {sum(r[1] for r in rows)} launches")
    for t, n, k in rows[:40]:
        print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}% n={n:5d} {k[:100]}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
