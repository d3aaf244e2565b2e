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
