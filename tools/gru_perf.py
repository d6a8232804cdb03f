"""Phase timing of the recurrent (GRU) MAPPO path at BASELINE configs[2] shape: simple_spread, 3 agents x 2048 envs,
T=25, shared actor-critic GRU, ppo_epoch 5, data_chunk_length 2 (examples/mpe/mpe_ppo.yaml).  Fast mode (device Philox)."""
import faulthandler, json, os, sys
faulthandler.dump_traceback_later(280, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from openrl_b200.configs.config import create_config_parser
from openrl_b200.envs.common import make
from openrl_b200.modules.common import PPONet
from openrl_b200.runners.common import PPOAgent
from openrl_b200.utils.logger import Logger

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = create_config_parser().parse_args(["--episode_length", "25", "--lr", "7e-4", "--critic_lr", "7e-4", "--use_recurrent_policy", "true",
                                         "--use_valuenorm", "true", "--use_adv_normalize", "true"])
cfg.quiet = True
env = make("simple_spread", env_num=N)
agent = PPOAgent(PPONet(env, cfg=cfg, device="cuda:0"))
agent.train(total_time_steps=0, logger=Logger(quiet=True))
drv = agent.driver
for _ in range(2):
    drv.device_iteration()
torch.cuda.synchronize()
drv.phase_events = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    drv.device_iteration()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
phases = {}
for name, a, b in drv.phase_events:
    phases[name] = phases.get(name, 0.0) + a.elapsed_time(b) / iters
info = drv.trainer.read_train_info()
print(json.dumps({"workload": f"simple_spread GRU MAPPO {N} envs x 3 agents, T=25, ppo_epoch {cfg.ppo_epoch}, L={cfg.data_chunk_length}",
                  "ms_per_iter": round(ms, 3), "env_steps_per_s": round(N * 25 / (ms * 1e-3)),
                  "phases_ms": {k: round(v, 3) for k, v in phases.items()}, "train_info": info}))
