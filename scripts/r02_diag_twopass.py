"""diagnostic: gradient reproducibility of two passes on ONE engine (default and small gradient buckets)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_trainer_gpu import _model, _batches, B, H, W
from dmmfods_b200.trainer import Trainer

def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

for bb in (32 << 20, 64 << 10):
    model = _model().cuda()
    tr = Trainer(model, B, H, W, bucket_bytes=bb)
    eng = tr.eng
    (x1, x2, tg), = _batches(1)
    gs, ls = [], []
    for i in range(3):
        eng.forward(x1, x2)
        eng.loss(tg)
        eng.backward()
        torch.cuda.synchronize()
        gs.append(eng.gflat.clone()); ls.append(eng.logits.clone())
    print("bucket_bytes %d: %d segments; pass 1 vs 0: grad %.3e logits %.3e; pass 2 vs 1: grad %.3e logits %.3e" % (
        bb, len(eng.segments), rel(gs[1], gs[0]), rel(ls[1], ls[0]), rel(gs[2], gs[1]), rel(ls[2], ls[1])))
    # per-parameter worst offenders
    worst = []
    for n in eng.param_names:
        o, k = eng.grad_offset[n], eng.p[n].numel()
        worst.append((rel(gs[1][o:o + k], gs[0][o:o + k]), gs[0][o:o + k].norm().item(), n))
    worst.sort(reverse=True)
    for w in worst[:6]:
        print("   %.3e  |g| %.3e  %s" % w)
