"""debug helper: per-stage relL2 of the engine's raw activations vs the bf16-emulated oracle."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dmmfods_b200 import config as cfgmod
from dmmfods_b200.model import Dense_U_Net_lidar
from oracle import dense_unet_oracle as du
from test_oracle_golden import load_tiny

name = sys.argv[1]
g, mc, sd, x1, x2, tgt = load_tiny(name)
c = cfgmod.get_config("/nonexistent")
for k, v in mc.items():
    setattr(c.model, k, v)
model = Dense_U_Net_lidar(c)
model.load_state_dict(sd)
model = model.cuda().train()
B, _, H, W = x1.shape
eng = model.engine(B, H, W)
out = eng.forward(x1.cuda(), x2.cuda())
torch.cuda.synchronize()
trace = {}
full = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
lg, _ = du.oracle_forward(full, mc, x1.double(), x2.double(), train=True, trace=trace, emulate_bf16=True)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def mat(m, C=None):
    C = C or m.ld
    return m.t[:, :C].float().cpu().reshape(m.B, m.H, m.W, C).permute(0, 3, 1, 2)


pairs = [("stem", "features.denseblock1", mc["num_init_features"])]
for b in range(4):
    pairs.append(("block%d" % (b + 1), "features.denseblock%d" % (b + 1), None))
for k in range(1, 5):
    pairs.append(("dec%d" % k, "decoder.%d" % k, None))
pairs.append(("refine0", "head.refine0", None))
for tn, en, C in pairs:
    t = trace[tn]
    m = eng.named[en]
    print("%-10s %-28s relL2 %.3e  max|ref| %.3g" % (tn, en, rel(mat(m, C or t.shape[1]), t), t.abs().max().item()))
print("logits relL2 %.3e" % rel(out.cpu(), lg))
