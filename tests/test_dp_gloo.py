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
