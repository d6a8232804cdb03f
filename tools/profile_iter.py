"""Small driver for ncu: a few device iterations of the bench workload (no e2e, no CPU baseline)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(300, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg, env, net, agent = bench.build_agent(0, 1)
drv = bench.make_driver(cfg, env, net, agent, 0, 1)
for _ in range(n):
    drv.device_iteration()
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == "e2e":
    from openrl_b200.utils.logger import Logger
    for k in (3, 10, 50):
        cfg2, env2, net2, agent2 = bench.build_agent(0, 1)
        cfg2.log_interval = 1
        torch.cuda.synchronize(); t0 = time.perf_counter()
        agent2.train(total_time_steps=bench.N_ENVS * bench.T * k, logger=Logger(quiet=True))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("e2e iters", k, "total ms", dt * 1e3, "ms/iter", dt * 1e3 / k)
        t0 = time.perf_counter()
        agent2.train(total_time_steps=bench.N_ENVS * bench.T * k, logger=Logger(quiet=True))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("  2nd call iters", k, "total ms", dt * 1e3, "ms/iter", dt * 1e3 / k)
