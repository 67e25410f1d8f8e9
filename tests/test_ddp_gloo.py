"""world_size-2 CPU (gloo) test of the multi-GPU host logic: every rank derives the SAME flat gradient
layout and bucket ranges from the launch plan, and the bucketed SUM all-reduce issued in backward-completion order
(trainer.BucketReducer, the object Trainer hands to Engine.backward) leaves sum_r grad_r on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmmfods_b200 import config as cfgmod


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dmmfods_b200.engine import Engine
        from dmmfods_b200.model import Dense_U_Net_lidar
        from dmmfods_b200.trainer import BucketReducer
        c = cfgmod.get_config("/nonexistent")
        c.model.stream_2_in_channels, c.model.concat_before_block_num = 1, 3
        c.model.growth_rate, c.model.block_config, c.model.num_init_features, c.model.bn_size = 16, (2, 4, 2, 2), 32, 2
        torch.manual_seed(123)
        m = Dense_U_Net_lidar(c)
        params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in m.state_dict(keep_vars=True).items()}
        eng = Engine(params, m.model_cfg(), 2, 64, 96, plan_only=True, bucket_bytes=64 << 10)
        # identical plan on every rank
        names = [None] * world
        dist.all_gather_object(names, (eng.param_names, [(s[3], s[4]) for s in eng.segments]))
        assert all(n == names[0] for n in names)
        # per-rank "gradients" (what the backward program would leave in the flat buffer)
        g = torch.Generator().manual_seed(1000 + rank)
        eng.gflat.copy_(torch.randn(eng.gflat.numel(), generator=g))
        mine = eng.gflat.clone()
        red = BucketReducer(dist, "cpu")
        for i, (_ops, _lo, _n, lo, hi) in enumerate(eng.segments):      # the order Engine.backward() calls on_bucket
            if hi > lo:
                red(i, eng.gflat[lo:hi])
        red.finish()
        allg = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allg, mine)
        want = sum(allg)
        assert torch.allclose(eng.gflat, want, rtol=0, atol=1e-6)
        assert sum(n for _, n in red.ranges) == eng.gflat.numel()
        # a named parameter's gradient view sees the reduced values (views of the flat buffer)
        k0 = eng.param_names[0]
        off = eng.grad_offset[k0]
        assert torch.equal(eng.grad[k0].flatten(), eng.gflat[off:off + eng.grad[k0].numel()])
        out.put((rank, "ok", len(eng.segments)))
    except Exception as e:      # noqa: BLE001
        out.put((rank, "fail: %r" % (e,), 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_bucketed_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res
    assert res[0][2] > 2      # several buckets
