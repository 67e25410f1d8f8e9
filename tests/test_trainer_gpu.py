"""-m gpu tests of the path whose throughput bench.py reports: `Trainer.step` (flat parameters, fused Adam, CUDA graph,
prefetch hand-over) against the LITERAL reference loop body (Dense_U_Net_lidar_Agent.py:244-265):

    prediction = model(image, lidar); current_loss = loss(prediction, ht_map)
    optimizer.zero_grad(); current_loss.backward(torch.ones_like(current_loss)); optimizer.step()

with `model` = the drop-in module, `loss` = FusedBCEWithLogits and `optimizer` = torch.optim.Adam, plus the checkpoint
round trip of Agent.py:96-163 (state_dict -> load_state_dict) and the batched-file reader feeding the trainer (N3).
Both arms run the same CUDA kernels; they differ in everything around them (flat buffers vs per-tensor parameters, fused
vs torch Adam, graph replay vs eager launches, prefetch ring vs direct copies).  The weight-gradient split-K uses fp32
reduce-adds whose order is not fixed, so gradients agree to fp32 round-off (1e-7), not bit for bit.  This small network at
64x96 (train-mode BatchNorm over 12 ... 3 000 samples, bf16 storage) is CHAOTIC: two runs of the SAME execution form stay
within 2e-9 in the parameters for two steps, then one flipped bf16 rounding is amplified to 4e-3 in the gradients of step 3 and
3e-1 in step 4 (scripts/r02_diag_trainer.py, measured on a B200); with other inputs the first flip already happens in step 2.
Parameter / BatchNorm-buffer equality is therefore asserted tightly after ONE optimisation step (round-off only: 1e-6) and
loosely after the second (1e-3: a doubled BatchNorm update or a wrong Adam step count would show as 1e-1).
"""
import copy
import os

import pytest
import torch

from dmmfods_b200 import config as cfgmod
from dmmfods_b200 import synthetic
from dmmfods_b200.data import BatchFileRing, list_batch_files
from dmmfods_b200.model import Dense_U_Net_lidar, FusedBCEWithLogits
from dmmfods_b200.trainer import Evaluator, StepLR, Trainer

pytestmark = pytest.mark.gpu

MC = {"growth_rate": 16, "block_config": (2, 2, 2, 2), "num_init_features": 32, "bn_size": 2,
      "stream_1_in_channels": 3, "stream_2_in_channels": 1, "concat_before_block_num": 3,
      "num_layers_before_blocks": 4, "drop_rate": 0, "num_classes": 3, "memory_efficient": False}
B, H, W = 2, 64, 96
LR = 1e-3


def _model(seed=123):
    c = cfgmod.get_config("/nonexistent")
    for k, v in MC.items():
        setattr(c.model, k, v)
    torch.manual_seed(seed)
    m = Dense_U_Net_lidar(c)
    gen = torch.Generator().manual_seed(5)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = torch.rand(mod.weight.shape, generator=gen) + 0.5
            mod.bias.data = torch.randn(mod.bias.shape, generator=gen) * 0.2
    return m


def _batches(n, pin=False):
    out = []
    for i in range(n):
        t = [torch.from_numpy(synthetic.rgb_image(B, H, W, seed=100 + i)), torch.from_numpy(synthetic.lidar_image(B, H, W, seed=200 + i)),
             torch.from_numpy(synthetic.target_maps(B, H, W, seed=300 + i))]
        out.append(tuple(v.pin_memory() if pin else v.cuda() for v in t))
    return out


def _agent_loop(model, batches, lr=LR):
    """Agent.py:244-265 literally (without logging): returns the per-class loss sums of every step."""
    loss_fn = FusedBCEWithLogits()
    opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False)
    sums = []
    model.train()
    for image, lidar, ht_map in batches:
        prediction = model(image.cuda(), lidar.cuda())
        current_loss = loss_fn(prediction, ht_map.cuda())
        sums.append(torch.sum(current_loss.detach(), dim=(0, 2, 3)).double().cpu())
        opt.zero_grad()
        current_loss.backward(torch.ones_like(current_loss.detach()))
        opt.step()
    torch.cuda.synchronize()
    return sums


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _compare_states(got, want, init, label, tol_p=1e-6, tol_u=1e-3, tol_bn=1e-5):
    """parameters: relL2 of the values and of the UPDATES (p - p_init); BatchNorm buffers; num_batches_tracked exact."""
    num = den = unum = uden = 0.0
    for k, v in want.items():
        g = got[k]
        if k.endswith("num_batches_tracked"):
            assert int(g) == int(v), (label, k, int(g), int(v))
            continue
        if "running_" in k:
            e = _rel(g.cpu(), v.cpu())
            assert e < tol_bn, "%s: %s relL2 %.3e" % (label, k, e)
            continue
        d = (g.double().cpu() - v.double().cpu())
        num += (d ** 2).sum().item()
        den += (v.double() ** 2).sum().item()
        u = (v.double().cpu() - init[k].double().cpu())
        unum += (d ** 2).sum().item()
        uden += (u ** 2).sum().item()
    e_p, e_u = (num / den) ** 0.5, (unum / uden) ** 0.5
    print("\n[%s] parameters after the steps: relL2 %.3e; relL2 of the updates %.3e" % (label, e_p, e_u))
    # Adam's first updates are lr * sign-like (g / |g|): an element whose gradient is pure round-off may move the other way
    assert e_p < tol_p, (label, e_p)
    assert e_u < tol_u, (label, e_u)


@pytest.mark.parametrize("use_graph,prefetch", [(False, False), (True, False), (True, True), (False, True)])
def test_trainer_steps_equal_the_reference_loop_body(use_graph, prefetch):
    K = 2
    ref_model = _model().cuda()
    init = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    ref_sums = _agent_loop(ref_model, _batches(1))
    want1 = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    ref_model = _model().cuda()
    ref_sums = _agent_loop(ref_model, _batches(K))
    want = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}

    model = _model().cuda()
    tr = Trainer(model, B, H, W, lr=LR, use_graph=use_graph)
    batches = _batches(K, pin=prefetch)
    sums = []
    for i, (x1, x2, tg) in enumerate(batches):
        nxt = batches[i + 1] if (prefetch and i + 1 < K) else None
        if prefetch and i == 0:
            tr.prefetch(x1, x2, tg)
        cs = tr.step(x1, x2, tg, prefetch_next=nxt)
        sums.append(cs.clone().cpu())
        if i == 0:
            torch.cuda.synchronize()
            got1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    torch.cuda.synchronize()
    assert _rel(sums[0], ref_sums[0]) < 2e-5, (sums[0], ref_sums[0])     # the reference loop sums the fp32 loss tensor in fp32, the engine in fp64
    assert _rel(sums[1], ref_sums[1]) < 1e-3, (sums[1], ref_sums[1])
    label = "Trainer(use_graph=%s, prefetch=%s)" % (use_graph, prefetch)
    _compare_states(got1, want1, init, label + ", one step")
    _compare_states(model.state_dict(), want, init, label + ", two steps", tol_p=1e-3, tol_u=1.0, tol_bn=1e-2)
    # the module's parameters ARE views of the trainer's flat buffer (an optimizer or checkpoint sees the trained values)
    p0 = next(model.parameters())
    assert tr.pflat.data_ptr() <= p0.data_ptr() < tr.pflat.data_ptr() + tr.pflat.numel() * 4


def test_gradient_accumulation_semantics_of_the_module_path():
    """p.grad adopted as a view of the engine's flat buffer must not break `backward(); backward()` accumulation."""
    model = _model().cuda().train()
    (x1, x2, tg), = _batches(1)
    loss_fn = FusedBCEWithLogits()
    ls = loss_fn(model(x1, x2), tg)
    ls.backward(torch.ones_like(ls))
    g1 = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    ls = loss_fn(model(x1, x2), tg)
    ls.backward(torch.ones_like(ls))            # accumulates: p.grad = g1 + g2, g2 ~= g1 (BN running stats do not enter train mode)
    for k, p in model.named_parameters():
        assert _rel(p.grad, 2 * g1[k]) < 1e-3 or g1[k].abs().max() < 1e-6, k


def test_checkpoint_round_trip_and_evaluator_after_trainer():
    """Agent.py:96-163: state_dict() -> torch.save -> load_state_dict() into a fresh model gives the same network; the
    Evaluator built on the trained module sees the trained weights (shared parameter storage)."""
    import io
    model = _model().cuda()
    tr = Trainer(model, B, H, W, lr=LR, use_graph=True)
    batches = _batches(3)
    for x1, x2, tg in batches:
        tr.step(x1, x2, tg)
    buf = io.BytesIO()
    torch.save({"state_dict": model.state_dict(), "steps": tr.steps}, buf)
    buf.seek(0)
    ck = torch.load(buf)
    fresh = _model(seed=999).cuda()
    fresh.load_state_dict(ck["state_dict"], strict=True)
    for (k, a), (_, b) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k
    x1, x2, tg = batches[0]
    ev_a = Evaluator(model, B, H, W)
    ev_b = Evaluator(fresh, B, H, W, use_graph=False)
    oa = ev_a.step(x1, x2, tg)
    ob = ev_b.step(x1, x2, tg)
    torch.cuda.synchronize()
    assert torch.equal(oa["logits"], ob["logits"])                 # eval mode: no atomics, bit-identical
    assert torch.equal(oa["loss_per_class"], ob["loss_per_class"])
    model.eval()
    with torch.no_grad():
        lg = model(x1, x2)
    assert torch.equal(lg, oa["logits"])


def test_step_lr_and_amsgrad():
    model = _model().cuda()
    with pytest.raises(NotImplementedError):
        Trainer(model, B, H, W, amsgrad=True)
    tr = Trainer(model, B, H, W, lr=1e-2)
    sched = StepLR(tr, step_size=2, gamma=0.1)
    ref = torch.optim.lr_scheduler.StepLR(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-2), step_size=2, gamma=0.1)
    for _ in range(5):
        sched.step()
        ref.optimizer.step()
        ref.step()
        assert abs(tr.lr - ref.get_last_lr()[0]) < 1e-12


def test_batch_file_ring_feeds_the_trainer_bit_identically(tmp_path):
    """N3: (B,7,H,W) batch files -> BatchFileRing (pinned ring, reader thread) -> Trainer.step(prefetch_next=): the engine's
    input buffers hold exactly the file contents at every step, and the training result equals direct feeding."""
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "train", "seg0"))
    K = 5
    files = []
    for i in range(K):
        x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=400 + i))
        x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=500 + i))
        tg = torch.from_numpy(synthetic.target_maps(B, H, W, seed=600 + i))
        full = torch.cat([x1, x2, tg], 1)
        torch.save(full, os.path.join(root, "train", "seg0", "batch_%d" % i))
        files.append(full)
    os.makedirs(os.path.join(root, "train", "seg0", "labels"))
    rel = list_batch_files(root, "train")
    order = [int(r.rsplit("_", 1)[1]) for r in rel]

    model = _model().cuda()
    tr = Trainer(model, B, H, W, lr=LR, use_graph=True)
    ring = iter(BatchFileRing(root, rel, depth=2))
    cur = next(ring)
    tr.prefetch(*cur)
    step = 0
    after2 = None
    while cur is not None:
        nxt = next(ring, None)
        tr.step(*cur, prefetch_next=nxt)
        full = files[order[step]].cuda()
        torch.cuda.synchronize()
        assert torch.equal(tr.eng.in1, full[:, :3]) and torch.equal(tr.eng.in2, full[:, 3:4])
        assert torch.equal(tr._static_target, full[:, 4:])
        cur = nxt
        step += 1
        if step == 1:
            after2 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert step == K

    direct = _model().cuda()
    td = Trainer(direct, B, H, W, lr=LR, use_graph=False)
    for j in order[:1]:
        full = files[j].cuda()
        td.step(full[:, :3].contiguous(), full[:, 3:4].contiguous(), full[:, 4:].contiguous())
    torch.cuda.synchronize()
    init = {k: v.detach().clone() for k, v in _model().state_dict().items()}
    _compare_states(after2, direct.state_dict(), init, "ring-fed vs direct, after one step")
