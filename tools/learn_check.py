import faulthandler, os, sys
faulthandler.dump_traceback_later(400, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from openrl_b200.configs.config import create_config_parser
from openrl_b200.envs.common import make
from openrl_b200.modules.common import PPONet
from openrl_b200.runners.common import PPOAgent
def run(seed, tc, steps=20000):
    cfg = create_config_parser().parse_args(["--seed", str(seed), "--use_tensor_cores", str(tc)])
    cfg.quiet = True
    env = make("CartPole-v1", env_num=9)
    agent = PPOAgent(PPONet(env, cfg=cfg, device="cuda:0"))
    agent.train(total_time_steps=steps)
    ev = make("CartPole-v1", env_num=64)
    obs, _ = ev.reset(seed=123)
    totals = np.zeros(64); fin = np.zeros(64, bool)
    for _ in range(500):
        a, _ = agent.act(obs, deterministic=True)
        obs, r, d, _ = ev.step(a)
        totals += r[:, 0, 0] * (~fin); fin |= d[:, 0]
        if fin.all(): break
    return totals.mean()
for tc in (False, True):
    print("tc", tc, [round(run(s, tc), 1) for s in range(6)], flush=True)
