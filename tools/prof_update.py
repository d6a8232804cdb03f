"""Run a few C2 iterations (CartPole-v1, 4096 envs, T=128, 4 epochs) for ncu captures of the update / rollout kernels.
    ncu --set full --import-source on --clock-control none -k regex:ppo_fwdbwd_tc -s 4 -c 1 -o gpurun_out/tc_r2 python tools/prof_update.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

extra = sys.argv[1:]
cfg, env, net, agent = bench.build_agent(0, 1, "c2", extra_flags=extra)
drv = bench.make_driver(cfg, env, net, agent, 0, 1)
for _ in range(3):
    drv.device_iteration()
torch.cuda.synchronize()
print("ok", drv.trainer.use_tensor_cores)
