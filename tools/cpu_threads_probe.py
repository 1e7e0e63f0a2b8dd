"""Which thread count runs the CPU oracle fastest on this host?  (bounded probe, T=M=512)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import transfoxl_oracle as orc
cfg = orc.make_cfg(12, 8, 512, 2048, 512, 512, False, -1, 729)
P = orc.init_params(cfg, 1111, 0.01)
res = {}
for nt in [8, 16, 32, 64, os.cpu_count()]:
    torch.set_num_threads(nt)
    opt = orc.AdamState(P); mems = [None]
    g = torch.Generator().manual_seed(1)
    ts = []
    for s in range(3):
        tok = torch.randint(2, 560, (513, 1), generator=g)
        t0 = time.time()
        _, _, mems, _ = orc.train_step(cfg, dict(P), opt, [(tok[:-1], tok[1:], torch.zeros(1, dtype=torch.bool))], mems, lr=1e-4)
        ts.append(time.time() - t0)
    res[nt] = ts
    print(nt, ["%.2f" % t for t in ts], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "cpu_threads.json"), "w"))
