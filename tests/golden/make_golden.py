"""Generates the committed golden vectors by executing the UNMODIFIED reference (/root/reference) through
oracle/ref_shim.py on deterministic synthetic inputs.  Run in the authoring container only:

    python tests/golden/make_golden.py

Outputs (small, committed):
    tests/golden/lidar_heatmap.npz   integer-scatter helpers (helper:233-305, 430-515)
    tests/golden/tiny_unet_<fusion>.npz  a small Dense_U_Net_lidar (growth 16, blocks (2,2,2,2)):
                                     state_dict, inputs, logits, loss, gradients (fp32 run + fp64 run)
    tests/golden/dn121_mid_probe.npz DenseNet-121 mid-fusion (cb=3), seed-123 init: logits / loss / gradient
                                     probes and per-tensor parameter checksums (state_dict too large to commit)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from dmmfods_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def lidar_heatmap():
    _, helper = ref_shim.ref_modules()
    out = {}
    cases = [("full", 1280, 1920, 30000), ("small", 64, 96, 400), ("edge", 40, 50, 0)]
    for name, H, W, n in cases:
        pts = synthetic.lidar_points(n, H, W, seed=7 + n)
        if name == "small":   # hand-made edge cases: borders, far out of range (python slice wrap), duplicates
            extra = np.array([[0, 0, 1.5], [W - 1, H - 1, 2.5], [W - 2, H - 2, 3.5], [W + 10, 5, 4.5], [5, H + 10, 5.5],
                              [-1, -1, 6.5], [-3, 10, 7.5], [10, -3, 8.5], [-30, -30, 9.5], [3.7, 4.2, 10.5],
                              [20, 20, 80.0], [20, 20, 25.0], [21, 21, 25.000002], [40, 30, 75.5], [41, 30, 0.0]],
                             dtype=np.float32)
            pts = np.concatenate([pts, extra], 0)
        img = helper.lidar_array_to_image_like_tensor(pts, shape=(1, H, W), kernel_size=5)
        out["%s_points" % name] = pts
        out["%s_shape" % name] = np.array([1, H, W])
        if name == "full":
            # full-size image is 9.8 MB: keep a checksum + a crop, and the pooled output
            out["full_img_crop"] = img[:, 600:700, 900:1100].numpy().copy()
            out["full_img_sum"] = np.array([img.double().sum().item(), (img.double() ** 2).sum().item()])
        else:
            out["%s_img" % name] = img.numpy().copy()
        if H >= 20:
            out["%s_pooled" % name] = helper.pool_lidar_tensor(img.clone()).numpy().copy()
        labels = synthetic.boxes(None if name == "full" else 12, H, W, seed=11 + n)
        if name == "small":
            labels["p0"] = {"type": 2, "x": 1, "y": 2, "width": 3, "height": 4}
            labels["p1"] = {"type": 2, "x": 10, "y": 12, "width": 4, "height": 3}
            labels["p2"] = {"type": 2, "x": 30, "y": 20, "width": 17, "height": 23}
            labels["p3"] = {"type": 2, "x": 32, "y": 25, "width": 9, "height": 11}
            labels["v0"] = {"type": 1, "x": 0, "y": 0, "width": W, "height": 1}
            labels["c0"] = {"type": 4, "x": W - 5, "y": H - 6, "width": 5, "height": 6}
            labels["u0"] = {"type": 3, "x": 0, "y": 0, "width": 5, "height": 5}
        maps = helper.create_ground_truth_maps(labels, width_img=W, height_img=H)
        boxes = np.array([[e["type"], e["x"], e["y"], e["width"], e["height"]] for e in labels.values()], dtype=np.int32)
        out["%s_boxes" % name] = boxes.reshape(-1, 5)
        if name == "full":
            out["full_maps_pooled"] = helper.maxpool_tensor(maps).numpy().copy()
            out["full_maps_sum"] = np.array([maps.double().sum().item(), (maps.double() ** 2).sum().item()])
        else:
            out["%s_maps" % name] = maps.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "lidar_heatmap.npz"), **out)
    print("lidar_heatmap.npz", {k: v.shape for k, v in out.items()})


def _train_step(model, x1, x2, tgt, dtype):
    model = model.to(dtype)
    model.train()
    model.zero_grad()
    logits = model(x1.to(dtype), x2.to(dtype))
    loss = torch.nn.BCEWithLogitsLoss(reduction="none")(logits, tgt.to(dtype))
    loss.backward(torch.ones_like(loss))
    return logits.detach(), loss.detach()


def _gerr(a_params, b_params):
    num = sum(((p.grad.double() - q.grad.double()) ** 2).sum().item() for p, q in zip(a_params, b_params))
    den = sum((q.grad.double() ** 2).sum().item() for q in b_params)
    return (num / den) ** 0.5


def tiny_unet(fusion, B=2, H=64, W=96, tag=None, store_sd=True):
    """golden for a small Dense_U_Net_lidar.  Stored: state_dict, inputs (small case) or their seeds + checksum
    (large case), fp64 logits / loss / gradients / updated BN buffers, and the ERROR LEVELS of the reference's own
    fp32 run and of its own torch.autocast(bfloat16) run against fp64 (the bf16 yard-stick)."""
    import copy
    model_mod, _ = ref_shim.ref_modules()
    c2, cb = {"no": (0, 1), "early": (1, 1), "mid": (1, 3)}[fusion]
    cfg = ref_shim.ref_config(stream_2_in_channels=c2, concat_before_block_num=cb, growth_rate=16,
                              block_config=(2, 2, 2, 2), num_init_features=32, bn_size=2)
    torch.manual_seed(123)
    model = model_mod.Dense_U_Net_lidar(cfg)
    # make the problem well conditioned: random BN affine parameters (at the gamma=1, beta=0 init the
    # gradient of every BN weight that feeds another BN is exactly zero in exact arithmetic - SURVEY App. D)
    g = torch.Generator().manual_seed(5)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.2
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=31))
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=32))
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W, seed=33))
    out = {"shape": np.array([B, H, W]), "seeds": np.array([31, 32, 33]),
           "input_checksum": np.array([x1.double().sum().item(), x2.double().sum().item(), tgt.double().sum().item()]),
           "model_cfg": np.array([16, 2, 2, 2, 2, 32, 2, c2, cb])}
    if store_sd:
        for k, v in model.state_dict().items():
            out["sd/" + k] = v.numpy()
    m64 = copy.deepcopy(model)
    logits64, loss64 = _train_step(m64, x1, x2, tgt, torch.float64)
    out["logits64"] = logits64.float().numpy()
    out["loss64_sum"] = np.array([loss64.sum().item()])
    out["loss64_per_class"] = loss64.sum(dim=(0, 2, 3)).numpy()
    for k, p in m64.named_parameters():
        out["grad64/" + k] = p.grad.float().numpy()
    for k, v in m64.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["new/" + k] = v.float().numpy() if v.is_floating_point() else v.numpy()
    # the reference's own fp32 run
    m32 = copy.deepcopy(model)
    logits32, loss32 = _train_step(m32, x1, x2, tgt, torch.float32)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    out["ref_fp32_err"] = np.array([rel(logits32, logits64), _gerr(list(m32.parameters()), list(m64.parameters()))])
    if H * W <= 64 * 96:
        out["logits32"] = logits32.numpy()
        out["loss32"] = loss32.numpy()
    # the reference's own bf16 yard-stick: CPU autocast(bfloat16) vs fp64
    mbf = copy.deepcopy(model)
    mbf.train()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lg = mbf(x1, x2)
    ls = torch.nn.BCEWithLogitsLoss(reduction="none")(lg.float(), tgt)
    ls.backward(torch.ones_like(ls))
    out["ref_bf16_autocast_err"] = np.array([rel(lg.detach().float(), logits64),
                                             _gerr(list(mbf.parameters()), list(m64.parameters()))])
    # eval-mode forward with the updated running stats (fp64)
    m64.eval()
    with torch.no_grad():
        out["eval_logits64"] = m64(x1.double(), x2.double()).float().numpy()
    name = "tiny_unet_%s.npz" % (tag or fusion)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, len(out), "arrays; params", sum(p.numel() for p in model.parameters()),
          "ref fp32 err", out["ref_fp32_err"], "ref bf16 err", out["ref_bf16_autocast_err"])


def dn121_probe():
    model_mod, _ = ref_shim.ref_modules()
    cfg = ref_shim.ref_config(stream_2_in_channels=1, concat_before_block_num=3)
    torch.manual_seed(123)
    model = model_mod.densenet121_u_lidar(pretrained=False, config=cfg)
    B, H, W = 1, 64, 96
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=41))
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=42))
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W, seed=43))
    out = {}
    names, sums, asums = [], [], []
    for k, v in model.state_dict().items():
        if v.is_floating_point():
            names.append(k)
            sums.append(v.double().sum().item())
            asums.append(v.double().abs().sum().item())
    out["param_names"] = np.array(names)
    out["param_sum"] = np.array(sums)
    out["param_abs_sum"] = np.array(asums)
    logits, loss = _train_step(model, x1, x2, tgt, torch.float32)
    out["logits32"] = logits.numpy()
    out["loss32_per_class"] = loss.double().sum(dim=(0, 2, 3)).numpy()
    gn = {k: p.grad.double().norm().item() for k, p in model.named_parameters()}
    out["grad_names"] = np.array(list(gn.keys()))
    out["grad_norm32"] = np.array(list(gn.values()))
    out["grad_refine1"] = dict(model.named_parameters())["dec_out_to_heat_maps.refine1.weight"].grad.numpy()
    out["num_params"] = np.array([model.num_params])
    np.savez_compressed(os.path.join(OUT, "dn121_mid_probe.npz"), **out)
    print("dn121_mid_probe.npz num_params", model.num_params)


def fullsize_yardstick(H=640, W=960):
    """BASELINE config 3 at FULL resolution, batch 1, the exact state / inputs of tests/test_fullsize_gpu.py (seed-123 init,
    BatchNorm affine parameters randomised with generator 7, input seeds 11 / 12 / 13), executed by the UNMODIFIED reference:
    fp64 run (golden loss sums, a logits crop, gradient norms) and the reference's OWN error levels against it when it runs
    in fp32 and under torch.autocast('cpu', bfloat16) - the bf16 yard-stick SURVEY 8(c)(5) prescribes."""
    import copy
    model_mod, _ = ref_shim.ref_modules()
    cfg = ref_shim.ref_config(stream_2_in_channels=1, concat_before_block_num=3)
    torch.manual_seed(123)
    model = model_mod.densenet121_u_lidar(pretrained=False, config=cfg)
    g = torch.Generator().manual_seed(7)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.2
    x1 = torch.from_numpy(synthetic.rgb_image(1, H, W, seed=11))
    x2 = torch.from_numpy(synthetic.lidar_image(1, H, W, seed=12))
    tgt = torch.from_numpy(synthetic.target_maps(1, H, W, seed=13))
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    out = {"shape": np.array([1, H, W]), "seeds": np.array([11, 12, 13]),
           "param_checksum": np.array([sum(v.double().sum().item() for v in model.state_dict().values() if v.is_floating_point())])}
    m64 = copy.deepcopy(model)
    logits64, loss64 = _train_step(m64, x1, x2, tgt, torch.float64)
    out["loss64_per_class"] = loss64.sum(dim=(0, 2, 3)).numpy()
    out["logits64_crop"] = logits64[:, :, 300:340, 400:480].float().numpy()
    out["logits64_norm"] = np.array([logits64.norm().item()])
    out["grad64_names"] = np.array([k for k, _ in m64.named_parameters()])
    out["grad64_norm"] = np.array([p.grad.norm().item() for _, p in m64.named_parameters()])
    out["grad64_refine1"] = dict(m64.named_parameters())["dec_out_to_heat_maps.refine1.weight"].grad.float().numpy()
    out["grad64_conv0"] = dict(m64.named_parameters())["features.conv0.weight"].grad.float().numpy()
    m32 = copy.deepcopy(model)
    logits32, _ = _train_step(m32, x1, x2, tgt, torch.float32)
    out["ref_fp32_err"] = np.array([rel(logits32, logits64), _gerr(list(m32.parameters()), list(m64.parameters()))])
    del m32
    mbf = copy.deepcopy(model)
    mbf.train()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lg = mbf(x1, x2)
    ls = torch.nn.BCEWithLogitsLoss(reduction="none")(lg.float(), tgt)
    ls.backward(torch.ones_like(ls))
    out["ref_bf16_autocast_err"] = np.array([rel(lg.detach().float(), logits64), _gerr(list(mbf.parameters()), list(m64.parameters()))])
    out["ref_bf16_autocast_loss_rel"] = np.array([abs(ls.double().sum().item() - loss64.sum().item()) / loss64.sum().item()])
    np.savez_compressed(os.path.join(OUT, "fullsize_yardstick.npz"), **out)
    print("fullsize_yardstick.npz: ref fp32 err", out["ref_fp32_err"], "ref bf16-autocast err", out["ref_bf16_autocast_err"],
          "loss rel", out["ref_bf16_autocast_loss_rel"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fullsize":
        fullsize_yardstick()
        sys.exit(0)
    lidar_heatmap()
    for f in ("no", "early", "mid"):
        tiny_unet(f)
    tiny_unet("mid", B=2, H=256, W=384, tag="mid_large", store_sd=False)     # BASELINE config-1 resolution
    dn121_probe()
