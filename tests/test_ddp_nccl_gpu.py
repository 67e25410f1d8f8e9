"""-m gpu, 2 ranks over NCCL (skipped with fewer than 2 GPUs): on-hardware correctness of the data-parallel path
("allreduce-SUM of per-rank grads == the gradient of the concatenated batch with per-shard BN statistics").

Every rank trains on its own shard; checked on both ranks:
  * the flat gradient buffer after backward + bucketed all-reduce equals the SUM over ranks of the gradients each rank
    computes alone on its shard (fp32 round-off: the weight-gradient split-K reduce-adds are unordered),
  * eager launches and per-segment graphs (and, with DMM_TEST_GRAPH_NCCL=1, the experimental single step graph with the NCCL
    all-reduces captured inside) leave the SAME parameters, and the parameters of the two ranks stay bit-identical,
  * BatchNorm running statistics stay per rank (plain nn.BatchNorm2d in the reference) and differ between the ranks.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B, H, W = 2, 64, 96


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dmmfods_b200 import synthetic
        from dmmfods_b200.trainer import Trainer
        from test_trainer_gpu import _model

        def shard(i):
            s = 1000 * rank + i
            return tuple(torch.from_numpy(f(B, H, W, seed=s + o)).cuda() for f, o in
                         ((synthetic.rgb_image, 1), (synthetic.lidar_image, 2), (synthetic.target_maps, 3)))

        # ---- gradient equality ----
        # this rank's own gradient is captured INSIDE the same backward pass, bucket by bucket, right before the bucket is handed
        # to the all-reduce (a second pass would not do: with NCCL kernels co-scheduled the order of the BatchNorm-statistics
        # atomics changes, a float mean moves by an ulp, a bf16 rounding flips and this small chaotic network amplifies it to
        # 1e-3 in the gradients - measured: 9.2e-4 between two passes of one rank, 0 without NCCL in flight)
        model = _model().cuda()
        tr = Trainer(model, B, H, W, lr=1e-3, distributed=True, use_graph=False, bucket_bytes=64 << 10)
        eng = tr.eng
        x1, x2, tg = shard(0)
        local = torch.zeros_like(eng.gflat)

        def hook(i, flat):
            lo = (flat.data_ptr() - eng.gflat.data_ptr()) // 4
            local[lo:lo + flat.numel()].copy_(flat)
            tr.reducer(i, flat)
        eng.forward(x1, x2)
        eng.loss(tg)
        eng.backward(on_bucket=hook)
        tr.reducer.finish()
        torch.cuda.synchronize()
        reduced = eng.gflat.clone()
        allg = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(allg, local)
        want = sum(a.double() for a in allg)
        err = ((reduced.double() - want).norm() / want.norm()).item()
        nb = len([s for s in eng.segments if s[4] > s[3]])
        assert local.abs().sum().item() > 0 and not torch.equal(allg[0], allg[1])
        assert err < 1e-6, "all-reduced gradient vs sum of per-rank gradients: relL2 %.3e" % err
        assert nb >= 2, "expected several gradient buckets, got %d" % nb
        assert sum(n for _, n in tr.reducer.ranges) == eng.gflat.numel()

        # ---- three execution forms leave the same parameters; ranks stay in lock-step ----
        finals = {}
        forms = [("eager", dict(use_graph=False)), ("segment_graphs", dict(use_graph=True, graph_nccl=False))]
        if os.environ.get("DMM_TEST_GRAPH_NCCL", "0") != "0":      # experimental form (NCCL captured inside the step graph): opt-in
            forms.append(("one_graph", dict(use_graph=True, graph_nccl=True)))
        for label, kw in forms:
            m = _model().cuda()
            t = Trainer(m, B, H, W, lr=1e-3, distributed=True, bucket_bytes=64 << 10, **kw)
            for i in range(1):
                t.step(*shard(i))
            torch.cuda.synchronize()
            if label == "one_graph":
                assert t.seg_graphs is None and t.graph is not None, "the NCCL all-reduces were not captured in the step graph"
            finals[label] = t.pflat.clone()
            other = [torch.empty_like(t.pflat) for _ in range(world)]
            dist.all_gather(other, t.pflat)
            assert torch.equal(other[0], other[1]), "%s: parameters of the two ranks diverged" % label
            rm = torch.cat([v.flatten() for k, v in m.state_dict().items() if k.endswith("running_mean")])
            rms = [torch.empty_like(rm) for _ in range(world)]
            dist.all_gather(rms, rm)
            assert not torch.equal(rms[0], rms[1]), "BatchNorm statistics must stay per rank"
            nbt = [int(v) for k, v in m.state_dict().items() if k.endswith("num_batches_tracked")]
            assert set(nbt) == {1}, "%s: num_batches_tracked %s after 1 step" % (label, sorted(set(nbt)))
        ref = finals["eager"].double()
        errs = {k: ((v.double() - ref).norm() / ref.norm()).item() for k, v in finals.items()}
        # the forms agree up to the chaos described above (measured 6e-4 after one Adam step: elements whose gradient is
        # round-off move by +-lr); a wrong bucket range or a missed all-reduce would show as O(1e-1)
        assert all(e < 5e-3 for e in errs.values()), errs
        out.put((rank, "ok", err, errs))
    except Exception as e:      # noqa: BLE001
        import traceback
        out.put((rank, "fail: %r\n%s" % (e, traceback.format_exc()), 0, {}))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_nccl_gradient_sum_and_graph_forms():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res
    print("\n[2-rank NCCL] reduced-vs-sum relL2 %.3e; parameters vs eager after 1 step: %s" % (res[0][2], res[0][3]))
