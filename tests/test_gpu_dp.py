"""Hardware test of data-parallel correctness: 2 ranks x B samples must train like 1 rank x 2B —
with the gradient exchange overlapped with backward and the optimizer overlapped with the next forward, for both the
all-reduce mode (GradientOverlap) and the ZeRO-1 mode (ShardedDataParallel).  Needs 2 GPUs (skipped otherwise):
    gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -q
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = dict(
    llm=dict(vocab_size=256, hidden_size=256, intermediate_size=704, num_hidden_layers=3, num_attention_heads=4,
             num_key_value_heads=2, rope_theta=1e6, rms_norm_eps=1e-6, hidden_act="silu", model_type="qwen2"),
    vision=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56,
                patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5))
B, L, R, STEPS = 4, 12, 4, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(step: int, lo: int, hi: int, dev):
    """Rows [lo, hi) of the global batch of `step` (the same rows whoever asks)."""
    g = torch.Generator().manual_seed(1000 + step)
    n = 2 * B
    ids = torch.randint(1, 256, (n, L), generator=g)
    ids[:, 1] = -200
    mask = torch.ones(n, L, dtype=torch.long)
    mask[1, 9:] = 0
    mask[6, 10:] = 0
    images = torch.randn(n, 3, 56, 56, generator=g)
    actions = torch.rand(n, 112, generator=g) * 2 - 1
    noise = torch.randn(n, R, 16, 7, generator=g)        # [sample, repeat]: rows of a sample stay together
    t = torch.randint(0, 100, (n, R), generator=g)
    drop = torch.rand(n, R, generator=g) < 0.2
    sl = slice(lo, hi)
    k = hi - lo
    # the model repeats the batch R times as [r0 rows..., r1 rows...] (cogact_arch: actions.repeat(R, 1, 1))
    rep = lambda x: x[sl].transpose(0, 1).reshape(R * k, *x.shape[2:]).to(dev)   # noqa: E731
    return dict(input_ids=ids[sl].to(dev), attention_mask=mask[sl].to(dev), images=images[sl].to(dev),
                actions=actions[sl].to(dev), repeated_diffusion_steps=R, noise=rep(noise), timesteps=rep(t),
                drop_mask=rep(drop))


def _model(dev):
    from dexbotic_b200.model import CogActConfig, CogACTForCausalLM
    from oracle.weights import seeded_state_dict
    c = CogActConfig(llm_config=CFG["llm"], mm_vision_tower=CFG["vision"], action_model_type="DiT-S", action_dim=7,
                     chunk_size=16)
    m = CogACTForCausalLM(c, device=dev)
    m.load_state_dict(seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 11))
    m.train()
    return m


def _train(model, dp, lo, hi, dev):
    losses, grads = [], None
    for step in range(STEPS):
        model.zero_grad()
        out = model(**_batch(step, lo, hi, dev))
        out.loss.backward()
        if dp is not None:
            dp.finish()
        if step == 0:
            torch.cuda.synchronize()
            grads = (model.store.grad_a.float().clone(), model.store.grad_b.clone())
        model.optimizer_step(base_lr=1e-3)
        losses.append(out.loss.item())
        if step == 0:
            sd1 = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    return losses, grads, {k: v.detach().float().cpu() for k, v in model.state_dict().items()}, sd1


def _worker(rank, world, port, mode, ret):
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    from dexbotic_b200.parallel import GradientOverlap, ShardedDataParallel
    from dexbotic_b200.params import ParamStore
    ParamStore.SYMMETRIC = mode == "zero1_ce"       # copy-engine transport: exchanged buffers in symmetric memory
    model = _model(dev)
    model.store.async_optimizer = True
    dp = (ShardedDataParallel(model.store) if mode.startswith("zero1")
          else GradientOverlap(model.store, bucket_bytes=1 << 16))
    if mode.startswith("zero1"):
        assert dp.ce == (mode == "zero1_ce")
    losses, grads, sd, sd1 = _train(model, dp, rank * B, (rank + 1) * B, dev)
    if mode.startswith("zero1"):     # 1/N of the moments per rank, and the shard pieces tile region A
        assert dp.exp_avg.numel() == model.store.n_a // world + model.store.n_b
        pieces = [None] * world
        dist.all_gather_object(pieces, dp.piece)
        cover = sorted(x for ps in pieces for x in ps)
        assert cover[0][0] == 0 and cover[-1][1] == model.store.n_a
        assert all(a[1] == b[0] for a, b in zip(cover[:-1], cover[1:]))
    # after the exchange, rank 0's gradient view: all-reduce -> whole buffer averaged; zero1 -> only its pieces
    ret[rank] = dict(losses=losses, sd=sd, sd1=sd1, grad_a=grads[0].cpu(), grad_b=grads[1].cpu(),
                     pieces=dp.piece if mode.startswith("zero1") else None)
    dist.destroy_process_group()


def _run_dp(mode):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), mode, ret), nprocs=2, join=True)
    return ret[0], ret[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_train_like_one_rank_with_twice_the_batch():
    dev = torch.device("cuda", 0)
    single = _model(dev)
    single.store.async_optimizer = True
    s_losses, s_grads, s_sd, _ = _train(single, None, 0, 2 * B, dev)
    del single
    torch.cuda.empty_cache()
    ar0, ar1 = _run_dp("allreduce")
    z0, z1 = _run_dp("zero1")
    c0, c1 = _run_dp("zero1_ce")
    # (1) the replicas stay identical, in every mode
    for a, b in ((ar0, ar1), (z0, z1), (c0, c1)):
        for k in a["sd"]:
            assert torch.equal(a["sd"][k], b["sd"][k]), k
    # (2) mean of the per-rank losses == the full-batch loss (equal sample counts), every step
    for step in range(STEPS):
        for r0, r1 in ((ar0, ar1), (z0, z1), (c0, c1)):
            mean = 0.5 * (r0["losses"][step] + r1["losses"][step])
            assert abs(mean - s_losses[step]) < 5e-3 * abs(s_losses[step]), (step, mean, s_losses[step])
    # (3) step-0 gradients: average of two bf16 half-batch gradients vs the bf16 full-batch gradient
    ga = ar0["grad_a"]
    rel = ((ga - s_grads[0].cpu()).norm() / s_grads[0].cpu().norm()).item()
    cos = torch.nn.functional.cosine_similarity(ga, s_grads[0].cpu(), dim=0).item()
    assert rel < 3e-2 and cos > 0.999, (rel, cos)
    relb = ((ar0["grad_b"] - s_grads[1].cpu()).norm() / s_grads[1].cpu().norm()).item()
    assert relb < 2e-2, relb
    # ZeRO-1 sees the same averaged gradient on the pieces it owns (N = 2: a + b in either order)
    for r in (z0, z1):
        for a, b in r["pieces"]:
            assert torch.equal(r["grad_a"][a:b], ga[a:b])
    # (4) the two NCCL exchange modes give the same weights after the first optimizer step: same averaged gradients,
    #     same AdamW arithmetic; only the clip coefficient's fp32 summation order differs (shard sums vs one sweep).
    #     (Later steps are not comparable element-wise: AdamW turns a 1-ulp gradient difference on a near-zero
    #     gradient — k_proj.bias has none at all — into a full +-lr move, and the runs drift apart chaotically.)
    #     Compared as updates: ||d_allreduce - d_zero1|| / ||d_allreduce|| per tensor.  k_proj.bias is skipped: softmax
    #     is shift-invariant in the keys, its true gradient is 0 and what is left is reduction-order noise that AdamW
    #     normalises to +-lr.
    init = {k: v.detach().float().cpu() for k, v in _model(dev).state_dict().items()}
    diffs = []
    for k in ar0["sd1"]:
        if "k_proj.bias" in k:
            continue
        d_a, d_z = ar0["sd1"][k] - init[k], z0["sd1"][k] - init[k]
        if d_a.norm() == 0:
            assert d_z.norm() == 0, k
            continue
        rel = ((d_a - d_z).norm() / d_a.norm()).item()
        if rel > 2e-2:
            diffs.append((k, round(rel, 5)))
    assert not diffs, diffs[:10]
    # the copy-engine transport averages in fp32 (one rounding) where NCCL rounds twice: same gradient to bf16 precision
    for r in (c0, c1):
        for a, b in r["pieces"]:
            d = (r["grad_a"][a:b] - ga[a:b]).norm() / ga[a:b].norm().clamp_min(1e-20)
            assert d < 1e-2, (a, b, d.item())
    # (5) and they track the single-GPU run: AdamW's first steps move every weight by ~lr, so compare the UPDATE
    #     direction (a sign flip on a near-zero gradient costs 2*lr on that element, bf16 rounding makes a few)
    bad = []
    for run, tag in ((ar0, "allreduce"), (c0, "zero1_ce")):
        for k in s_sd:
            if "k_proj.bias" in k:          # no true gradient (see above)
                continue
            d_s = (s_sd[k] - init[k]).flatten()
            d_p = (run["sd"][k] - init[k]).flatten()
            if d_s.norm() == 0:
                continue
            c = torch.nn.functional.cosine_similarity(d_s, d_p, dim=0).item()
            if c < 0.97:
                bad.append((tag, k, round(c, 4)))
    assert not bad, bad[:8]
