"""CPU, world_size=2, gloo: the N>1 host logic (flat-buffer gradient averaging + parameter broadcast)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dexbotic_b200.parallel import allreduce_gradients, broadcast_parameters
    from dexbotic_b200.params import ParamSpec, ParamStore
    specs = [ParamSpec("a.weight", (8, 16)), ParamSpec("a.bias", (8,)), ParamSpec("h.weight", (4, 8), compute="fp32"),
             ParamSpec("frozen.weight", (3, 8), trainable=False)]
    st = ParamStore(specs, "cpu")
    st.master.copy_(torch.arange(st.master.numel(), dtype=torch.float32) * (rank + 1))
    broadcast_parameters(st, 0)
    st.g("a.weight").fill_(float(rank + 1))
    st.g("a.bias").fill_(2.0 * (rank + 1))
    st.g("h.weight").fill_(10.0 * (rank + 1))
    allreduce_gradients(st, bucket_elems=64)
    ok = (torch.equal(st.master, torch.arange(st.master.numel(), dtype=torch.float32))
          and bool((st.g("a.weight") == 1.5).all()) and bool((st.g("a.bias") == 3.0).all())
          and bool((st.g("h.weight") == 15.0).all()) and st.g("frozen.weight") is None)
    ret[rank] = ok
    dist.destroy_process_group()


def test_gradient_average_and_broadcast_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_param_store_layout_cpu():
    from dexbotic_b200.params import ParamSpec, ParamStore
    specs = [ParamSpec("l.q.weight", (8, 16), fuse="qkv"), ParamSpec("l.k.weight", (4, 16), fuse="qkv"),
             ParamSpec("l.v.weight", (4, 16), fuse="qkv"), ParamSpec("l.n.weight", (16,)),
             ParamSpec("head.weight", (5, 16), compute="fp32"), ParamSpec("lm_head.weight", (32, 16), trainable=False)]
    st = ParamStore(specs, "cpu")
    fused = st.fused_w(["l.q.weight", "l.k.weight", "l.v.weight"])
    assert fused.shape == (16, 16) and fused.data_ptr() == st.w("l.q.weight").data_ptr()
    assert st.fused_g(["l.q.weight", "l.k.weight", "l.v.weight"]).dtype == torch.bfloat16
    assert st.g("head.weight").dtype == torch.float32 and st.g("lm_head.weight") is None
    assert st.w("head.weight").data_ptr() == st.master_view("head.weight").data_ptr()      # fp32 compute = master
    segs = st.segments({"llm": 1e-3, "action_head": 1e-4}, 0.1)
    assert all(b <= st.n_train for _, b, *_ in segs)
    g = st.g("l.n.weight")
    assert st.first_write(g) and not st.first_write(g)
    st.zero_grad()
    assert st.first_write(g)


def test_optimizer_chunk_split_covers_every_element_once():
    """ParamStore.adamw_step's overlapped path: segments split at the per-block chunk boundaries must tile the
    trainable range exactly (no parameter updated twice or skipped), for sorted, interleaved (pi0: LLM / expert
    layers alternate) and partly frozen (None) chunk lists, and for segments that alternate weight decay."""
    import random
    from dexbotic_b200.params import split_segments
    rnd = random.Random(0)
    for trial in range(200):
        n = rnd.randrange(2000, 6000)
        # segments: consecutive runs over [0, n) in region A with alternating (lr, wd), then one region-B segment
        cuts = sorted(rnd.sample(range(64, n, 64), rnd.randrange(1, 8)))
        segs, prev = [], 0
        for j, c in enumerate(cuts + [n]):
            segs.append((prev, c, 1e-3 if j % 3 else 2e-3, 0.0 if j % 2 else 0.1, "A"))
            prev = c
        segs.append((n, n + 500, 1e-4, 0.0, "B"))
        # chunks: disjoint ranges, shuffled order, some None
        pts = sorted(rnd.sample(range(0, n, 32), 2 * rnd.randrange(1, 6)))
        chunks = [(pts[i], pts[i + 1]) for i in range(0, len(pts), 2)]
        rnd.shuffle(chunks)
        chunks = [c if rnd.random() > 0.2 else None for c in chunks]
        rest, per_chunk = split_segments(segs, chunks)
        assert len(per_chunk) == len(chunks)
        cover = [0] * (n + 500)
        for piece in rest + [p for ps in per_chunk for p in ps]:
            a, b, lr, wd, region = piece
            assert a < b
            src = [sg for sg in segs if sg[0] <= a and b <= sg[1] and sg[4] == region]
            assert len(src) == 1 and (lr, wd) == (src[0][2], src[0][3]), "a piece keeps its segment's lr / wd"
            for e in range(a, b):
                cover[e] += 1
        assert all(c == 1 for c in cover), (trial, [i for i, c in enumerate(cover) if c != 1][:5])
        for i, c in enumerate(chunks):
            for a, b, *_ in per_chunk[i]:
                assert c is not None and c[0] <= a and b <= c[1], "chunk pieces stay inside their chunk"


# ------------------------------------------------------------------------------------------- ZeRO-1 host logic
def _torch_adamw(p, g, m, v, shadow, lr, b1, b2, eps, wd, step, clip=None):
    """torch.optim.AdamW's single-tensor arithmetic (stand-in for the CUDA kernel in the CPU test)."""
    gf = g.float() * (1.0 if clip is None else float(clip))
    m.lerp_(gf, 1 - b1)
    v.mul_(b2).addcmul_(gf, gf, value=1 - b2)
    p.mul_(1 - lr * wd)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)
    if shadow is not None:
        shadow.copy_(p.to(shadow.dtype))


def _torch_sumsq(x, out):
    out += x.float().pow(2).sum()


def _torch_clip(ssq, max_norm, clip, norm_out=None):
    n = ssq.sqrt()
    clip.copy_(torch.clamp(max_norm / (n + 1e-6), max=1.0))
    if norm_out is not None:
        norm_out.copy_(n)


def _zero1_specs():
    from dexbotic_b200.params import ParamSpec
    sp = [ParamSpec("embed.weight", (40, 16)), ParamSpec("tower.fc.weight", (24, 16), group="vision"),
          ParamSpec("tower.fc.bias", (24,), group="vision")]
    for i in range(3):
        q = f"layers.{i}."
        sp += [ParamSpec(q + "q.weight", (16, 16), fuse=f"qkv{i}"), ParamSpec(q + "k.weight", (8, 16), fuse=f"qkv{i}"),
               ParamSpec(q + "v.weight", (8, 16), fuse=f"qkv{i}"), ParamSpec(q + "q.bias", (16,)),
               ParamSpec(q + "norm.weight", (16,)), ParamSpec(q + "mlp.weight", (48, 16))]
    sp += [ParamSpec("head.weight", (7, 16), group="action_head", compute="fp32"),
           ParamSpec("head.bias", (7,), group="action_head", compute="fp32"),
           ParamSpec("frozen.weight", (5, 16), trainable=False)]
    return sp


def _zero1_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dexbotic_b200.parallel import ShardedDataParallel
    from dexbotic_b200.params import ParamStore
    lrs = {"llm": 1e-2, "vision": 3e-3, "action_head": 2e-2}

    def fresh():
        st = ParamStore(_zero1_specs(), "cpu")
        g = torch.Generator().manual_seed(5)
        st.master[: st.n_train].copy_(torch.randn(st.n_train, generator=g) * 0.1)
        st.shadow.copy_(st.master[: st.n_a].to(torch.bfloat16))
        st.set_param_chunks([st.grad_range([n for n in st.order if n.startswith(f"layers.{i}.")]) for i in range(3)])
        return st

    def local_grads(st, r, step):
        g = torch.Generator().manual_seed(100 * step + r)
        ga = (torch.randn(st.n_a, generator=g) * (1.0 + r)).to(torch.bfloat16)
        gb = torch.randn(st.n_b, generator=g) * (1.0 + r)
        return ga, gb

    st = fresh()
    dp = ShardedDataParallel(st, adamw_fn=_torch_adamw, sumsq_fn=_torch_sumsq, clip_fn=_torch_clip)
    assert dp.enabled and st.sharder is dp
    # chunks tile region A, pieces tile the chunks across ranks
    assert dp.chunks[0][0] == 0 and dp.chunks[-1][1] == st.n_a
    assert all(a[1] == b[0] for a, b in zip(dp.chunks[:-1], dp.chunks[1:]))
    assert len([c for c in dp.block_chunk if c is not None]) == 3 and len(dp.rest) == len(dp.chunks) - 3
    # reference: both ranks' gradients averaged, ONE unsharded optimizer (same stand-in kernels)
    ref = fresh()
    ref_m = torch.zeros(ref.n_train)
    ref_v = torch.zeros(ref.n_train)
    ok = True
    for step in (1, 2, 3):
        st.zero_grad()
        ga, gb = local_grads(st, rank, step)
        st.grad_a.copy_(ga)
        st.grad_b.copy_(gb)
        st._written_ranges.append((0, st.n_a))          # every tensor was written this step
        # block backward order: last layer first; the hook reduce-scatters that block's chunk
        for i in (2, 1, 0):
            st.grad_ready_hook(*st._chunk_bounds[i])
        assert len(dp.reduced) == 3
        dp.finish()
        norm = st.adamw_step(lrs, weight_decay=0.05, max_grad_norm=1.0)
        # ---- reference
        gas = [local_grads(ref, r, step) for r in range(world)]
        # averaged in fp32 and rounded to bf16 once, exactly as the gloo stand-in of the reduce-scatter does
        ra = (sum(x[0].float() for x in gas) / world).to(torch.bfloat16)
        rb = sum(x[1] for x in gas) / world
        ssq = ra.float().pow(2).sum() + rb.pow(2).sum()
        clip = torch.clamp(1.0 / (ssq.sqrt() + 1e-6), max=1.0)
        for a, b, lr, wd, region in ref.segments(lrs, 0.05):
            g = ra[a:b] if region == "A" else rb[a - ref.n_a:b - ref.n_a]
            sh = ref.shadow[a:b] if region == "A" else None
            _torch_adamw(ref.master[a:b], g, ref_m[a:b], ref_v[a:b], sh, lr, 0.9, 0.999, 1e-8, wd, step, clip)
        ok &= bool(torch.allclose(norm, ssq.sqrt(), rtol=1e-5))
        ok &= bool(torch.equal(st.shadow, ref.shadow))                          # gathered compute copy: identical
        pa = torch.cat([st.master[a:b] for a, b in dp.piece])
        ok &= bool(torch.allclose(pa, torch.cat([ref.master[a:b] for a, b in dp.piece]), rtol=0, atol=1e-7))
        ok &= bool(torch.allclose(st.master[st.n_a:st.n_train], ref.master[ref.n_a:ref.n_train], atol=1e-7))
    dp.gather_master()
    ok &= bool(torch.allclose(st.master[: st.n_train], ref.master[: ref.n_train], rtol=0, atol=1e-7))
    ok &= dp.exp_avg.numel() == st.n_a // world + st.n_b                         # 1/N of the moments per rank
    # gradient accumulation: nothing is exchanged under no_sync()
    st.zero_grad()
    with dp.no_sync():
        st.grad_ready_hook(*st._chunk_bounds[0])
        dp.finish()
    ok &= len(dp.reduced) == 0
    ret[rank] = ok
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 4])
def test_zero1_sharded_step_equals_unsharded(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_zero1_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _overlap_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dexbotic_b200.parallel import GradientOverlap
    from dexbotic_b200.params import ParamStore
    st = ParamStore(_zero1_specs(), "cpu")
    ov = GradientOverlap(st, bucket_bytes=1)
    # gloo cannot all-reduce bf16 with AVG: exercise the bookkeeping only (what is launched when)
    launched = []
    ov._launch = lambda a, b: (launched.append((a, b)), ov.done.append((a, b)))
    st.set_param_chunks([st.grad_range([n for n in st.order if n.startswith(f"layers.{i}.")]) for i in range(3)])
    st.zero_grad()
    with ov.no_sync():                                   # micro-batch 1 of 2: gradients stay local
        for i in (2, 1, 0):
            st.grad_ready_hook(*st._chunk_bounds[i])
        ov.finish()
    ok = launched == []
    for i in (2, 1, 0):                                  # last micro-batch: every block range reduced exactly once
        st.grad_ready_hook(*st._chunk_bounds[i])
    ok &= sorted(launched) == sorted(st._chunk_bounds)
    st.zero_grad()                                       # a new step forgets what was reduced
    ok &= ov.done == [] and ov.pending is None
    ret[rank] = ok
    dist.destroy_process_group()


def test_gradient_overlap_accumulation_bookkeeping_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_overlap_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
