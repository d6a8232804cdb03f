"""Cycle attribution of the tcgen05 update kernel's tile pipeline (thread 0 of every CTA accumulates
clock() deltas per phase into spare slots of its partial row)."""
import faulthandler, os, sys
faulthandler.dump_traceback_later(200, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
cfg, env, net, agent = bench.build_agent(0, 1)
drv = bench.make_driver(cfg, env, net, agent, 0, 1)
for _ in range(3):
    drv.device_iteration()
torch.cuda.synchronize()
tr = drv.trainer
P = tr.partials.cpu().numpy()
G = tr.grid_per_net
names = ["gather+fc1+LN1 -> GEMM1 issue", "wait GEMM1", "LN3+head+loss+dZ3 -> GEMM2 issue", "wait GEMM2", "LN1-bwd, GEMM3 issue, GH"]
for net_name, rows in (("policy", P[:G]), ("critic", P[G:])):
    prof = rows[:, -5:]
    tot = prof.sum(axis=1).mean()
    tiles = (bench.N_ENVS * bench.T + 127) // 128 / G
    print(net_name, "cycles per CTA %.0f, per tile %.0f" % (tot, tot / tiles))
    for i, nm in enumerate(names):
        print("   %-36s %7.0f cycles/tile  %5.1f%%" % (nm, prof[:, i].mean() / tiles, prof[:, i].mean() / tot * 100))
