"""How much of a bench step is host launch overhead?  Times one workload three ways on one GPU:
  eager/async   the bench default (AdamW of step t on a side stream under the forward of step t+1)
  eager/sync    the same launches with the optimizer in line
  graph/sync    zero_grad + forward + backward replayed as ONE CUDA graph, optimizer eager and in line
Usage: python tools/graph_probe.py [workload] [steps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cogact_7b"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    w = bench.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = bench.build_model(w, dev)
    model.init_weights_(seed=1234)
    model.train()
    batch = {k: (v.to(dev) if hasattr(v, "to") else v) for k, v in bench.make_batch(w, 0, pinned=False).items()}
    if w.get("kind") != "pi0":
        P = (w["vision"]["image_size"] // w["vision"]["patch_size"]) ** 2
        extra = w.get("extra_tokens", 0) if w.get("kind") == "oft_l1" else 0
        model.config.static_seq_len = batch["input_ids"].shape[1] - 1 + P + extra

    def fb():
        model.zero_grad()
        out = model(**batch)
        out.loss.backward()
        return out.loss

    def timeit(fn, n):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        model.store.wait_all_params()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    def eager_step():
        fb()
        model.optimizer_step(base_lr=2e-5)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):            # everything on ONE side stream (autograd caches the leaf streams)
        model.store.async_optimizer = True
        for _ in range(3):
            eager_step()
        t_async = timeit(eager_step, steps)
        model.store.async_optimizer = False
        eager_step()
        t_sync = timeit(eager_step, steps)
        t_fb = timeit(fb, steps)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph, stream=side):
                loss = fb()
        except Exception as e:                  # a host sync on the path: say where
            import traceback
            traceback.print_exc()
            print(f"{name}: eager/async {t_async:.2f}  eager/sync {t_sync:.2f}  fwd+bwd {t_fb:.2f} ms; capture failed: {e}")
            return

        def graph_step():
            graph.replay()
            model.optimizer_step(base_lr=2e-5)

        graph_step()
        t_graph = timeit(graph_step, steps)
        t_gfb = timeit(graph.replay, steps)
    B = w["batch"]
    print(f"{name}: B={B}  loss {loss.item():.4f}")
    for k, v in (("eager/async (bench default)", t_async), ("eager/sync", t_sync), ("graph(fwd+bwd)/sync opt", t_graph),
                 ("eager fwd+bwd only", t_fb), ("graph fwd+bwd only", t_gfb)):
        print(f"  {k:30s} {v:8.2f} ms/step  {B / v * 1e3:8.2f} samples/s")


if __name__ == "__main__":
    main()
