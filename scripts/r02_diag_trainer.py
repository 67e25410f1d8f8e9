"""diagnostic: per-step divergence between Trainer execution forms (eager / graph / prefetch) on the tiny test network."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_trainer_gpu import _model, _batches, B, H, W, LR
from dmmfods_b200.trainer import Trainer

K = 5
def run(use_graph, prefetch):
    model = _model().cuda()
    tr = Trainer(model, B, H, W, lr=LR, use_graph=use_graph)
    batches = _batches(K, pin=prefetch)
    out = []
    for i, (x1, x2, tg) in enumerate(batches):
        nxt = batches[i + 1] if (prefetch and i + 1 < K) else None
        if prefetch and i == 0:
            tr.prefetch(x1, x2, tg)
        cs = tr.step(x1, x2, tg, prefetch_next=nxt)
        torch.cuda.synchronize()
        sd = model.state_dict()
        out.append(dict(cs=cs.clone().cpu(), p=tr.pflat.clone().cpu(), g=tr.eng.gflat.clone().cpu(),
                        rm=sd["features.norm0.running_mean"].clone().cpu(), rv=sd["features.norm0.running_var"].clone().cpu(),
                        rm2=sd["features.denseblock2.denselayer1.norm1.running_mean"].clone().cpu(),
                        nbt=int(sd["features.norm0.num_batches_tracked"])))
    return out

def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

runs = {"eager": run(False, False), "eager2": run(False, False), "graph": run(True, False), "graph+prefetch": run(True, True), "eager+prefetch": run(False, True)}
ref = runs["eager"]
for name, r in runs.items():
    for i in range(K):
        print("%-15s step %d: cs %.2e  grad %.2e  params %.2e  rm0 %.2e  rv0 %.2e  rm(b2l1) %.2e  nbt %d" % (
            name, i, rel(r[i]["cs"], ref[i]["cs"]), rel(r[i]["g"], ref[i]["g"]), rel(r[i]["p"], ref[i]["p"]), rel(r[i]["rm"], ref[i]["rm"]),
            rel(r[i]["rv"], ref[i]["rv"]), rel(r[i]["rm2"], ref[i]["rm2"]), r[i]["nbt"]))
