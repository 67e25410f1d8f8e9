"""smoke test used by __graft_entry__.smoke(): one small training step of the hot path on cuda:0 (forward, BCE loss,
backward through the C-ABI kernels) + the integer scatters, checked against the CPU oracle."""
import os
import sys

import numpy as np
import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from dmmfods_b200 import config as cfgmod, helper, synthetic
    from dmmfods_b200.model import Dense_U_Net_lidar, FusedBCEWithLogits
    from oracle import dense_unet_oracle as du
    from oracle import lidar_heatmap_oracle as orc

    torch.cuda.set_device(0)
    mc = {"growth_rate": 16, "block_config": (2, 2, 2, 2), "num_init_features": 32, "bn_size": 2,
          "stream_1_in_channels": 3, "stream_2_in_channels": 1, "concat_before_block_num": 3,
          "num_layers_before_blocks": 4, "drop_rate": 0, "num_classes": 3, "memory_efficient": False}
    c = cfgmod.get_config("/nonexistent")
    for k, v in mc.items():
        setattr(c.model, k, v)
    torch.manual_seed(123)
    model = Dense_U_Net_lidar(c)
    gen = torch.Generator().manual_seed(5)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=gen) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.2
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    B, H, W = 2, 128, 192
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=1))
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=2))
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W, seed=3))
    model = model.cuda().train()
    logits = model(x1.cuda(), x2.cuda())
    loss = FusedBCEWithLogits()(logits, tgt.cuda())
    loss.backward(torch.ones_like(loss))
    torch.cuda.synchronize()
    ref = du.oracle_train_step(sd, mc, x1, x2, tgt, dtype=torch.float64, emulate_bf16=True)

    def rel(a, b):
        return ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
    e_logits = rel(logits.detach(), ref["logits"])
    num = sum(((p.grad.double().cpu() - ref["grads"][k]) ** 2).sum().item() for k, p in model.named_parameters())
    den = sum((v ** 2).sum().item() for v in ref["grads"].values())
    e_grad = (num / den) ** 0.5
    e_loss = abs(loss.double().sum().item() - ref["loss"].sum().item()) / ref["loss"].sum().item()
    print("smoke: Dense-U-Net mid-fusion step on %s: logits relL2 %.3e, loss rel %.3e, grad relL2 %.3e (vs bf16-emulated oracle)"
          % (torch.cuda.get_device_name(0), e_logits, e_loss, e_grad))
    assert e_logits < 4e-2 and e_loss < 1e-3 and e_grad < 1.5e-1, "CUDA hot path disagrees with the oracle"

    pts = synthetic.lidar_points(2000, 200, 300, seed=4, out_of_range=0.05)
    img = helper.lidar_array_to_image_like_tensor(pts, shape=(1, 200, 300))
    assert np.array_equal(img.cpu().numpy(), orc.lidar_array_to_image(pts, (1, 200, 300), 5)), "lidar splat mismatch"
    assert np.array_equal(helper.pool_lidar_tensor(img).cpu().numpy(), orc.pool_lidar(img.cpu().numpy())), "lidar pool mismatch"
    labels = synthetic.boxes(20, 200, 300, seed=5)
    maps = helper.create_ground_truth_maps(labels, width_img=300, height_img=200)
    assert np.array_equal(maps.cpu().numpy(), orc.create_ground_truth_maps(labels, 300, 200)), "heat-map mismatch"
    print("smoke: lidar splat / pool / heat-map masks bit-exact vs oracle")
