import faulthandler; faulthandler.dump_traceback_later(90, exit=True)
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_ppo_update_cuda import _setup, _load_buffer
from oracle import loop, ppo as oppo
d = np.load(os.path.join(ROOT, "tests/golden/trace_cartpole.npz"), allow_pickle=True)
cfg, net, trainer, buf = _setup(d)
_load_buffer(buf, d, 0)
vn = net.module.get_critic_value_normalizer()
buf.data.compute_returns(buf.data.value_preds[-1].clone(), vn)
total = cfg.episode_length * int(d["meta/env_num"])
mb = total // cfg.num_mini_batch
perm = torch.from_numpy(d["it0/perms"][0]).cuda()
idx = perm[:mb].contiguous()
ocfg = loop.cfg_from_flags(str(d["meta/flags"]))
pol = {k: torch.from_numpy(d["init/policy." + k]).clone() for k, _ in net.module.models["policy"].named_parameters()}
cri = {k: torch.from_numpy(d["init/critic." + k]).clone() for k, _ in net.module.models["critic"].named_parameters()}
opt_p, opt_c = oppo.make_optimizers(ocfg, pol, cri)
ovn = oppo.ValueNormState()
flat = lambda x: torch.from_numpy(x.reshape(total, -1))
ii = torch.from_numpy(d["it0/perms"][0][:mb])
batch = dict(critic_obs=flat(d["it0/policy_obs"][:-1])[ii], policy_obs=flat(d["it0/policy_obs"][:-1])[ii],
             actions=flat(d["it0/actions"])[ii], value_preds=flat(d["it0/value_preds"][:-1])[ii],
             returns=flat(d["it0/returns"][:-1])[ii], active_masks=flat(d["it0/active_masks"][:-1])[ii],
             old_logp=flat(d["it0/action_log_probs"])[ii], adv=flat(d["it0/advantages"])[ii],
             action_masks=flat(d["it0/action_masks"][:-1])[ii])
ocfg.use_max_grad_norm = False
res = oppo.ppo_update(ocfg, pol, cri, opt_p, opt_c, ovn, batch)
print("oracle", res)
trainer.lrs.copy_(torch.tensor([cfg.lr, cfg.critic_lr]))
trainer.train_info.zero_()
trainer.ppo_update(buf.data, mb, idx)
torch.cuda.synchronize()
print("cuda  ", trainer.train_info.cpu().numpy())
print("gae_stats", buf.data.gae_stats.cpu().numpy(), "mb_stats", trainer.mb_stats.cpu().numpy())
print("ret mb sum", batch["returns"].double().sum().item(), (batch["returns"].double()**2).sum().item(), batch["active_masks"].sum().item())
print("vn cuda", vn.state.cpu().numpy(), "vn oracle", ovn.state())
grads = trainer.grads.cpu().numpy()
for net_i, params in ((0, pol), (1, cri)):
    off = 0
    for k, p in params.items():
        n = p.numel()
        want = p.grad.numpy().reshape(-1); got = grads[net_i, off:off+n]
        err = np.abs(got-want).max(); print(net_i, k, "max|w|=%.3e max|g|=%.3e err=%.3e" % (np.abs(want).max(), np.abs(got).max(), err))
        off += n
adv_dev = buf.data.advantages.cpu().numpy().reshape(-1)
print("raw adv vs golden-normalised: mean/std", adv_dev.mean(), adv_dev.std())
