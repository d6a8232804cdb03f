"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the two
exchange steps of the path:

  1. once per iteration: SUM all-reduce of the 8 float64 rollout moments (advantage / return sums
     and counts, `orl_gae` stats) so that advantage normalisation (ppo.py:402-409), masked-mean
     denominators (ppo.py:213-217,315-317) and ValueNorm batch moments (valuenorm.py:64-65) are the
     global-batch values;
  2. once per update: SUM all-reduce of the folded gradient bucket `folded` (2*stride floats,
     38 KB for CartPole) between orl_ppo_reduce and orl_ppo_apply.

Rollouts shard over envs: rank r owns global envs [r*N, (r+1)*N) and seeds them exactly as the
unsharded vec-env would (seed + global_index*10086, sync_venv.py:137); no exchange during
collection or GAE.  The reference has no working distributed path: this is new
design.  The same functions run on CPU tensors with the gloo backend (tests/test_multiproc_cpu.py).
"""
import torch.distributed as dist


_FORCE_SINGLE = False


def force_single_process(on=True):
    """Make this process behave as an unsharded run even though a process group exists (used by the multi-GPU
    equivalence check, which runs the global-batch reference on rank 0 next to the sharded run)."""
    global _FORCE_SINGLE
    _FORCE_SINGLE = bool(on)


def is_distributed():
    return (not _FORCE_SINGLE) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if is_distributed() else 1


def rank():
    return dist.get_rank() if is_distributed() else 0


def allreduce_sum_(t):
    """In-place SUM all-reduce (no-op for a single process)."""
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def env_shard(global_envs, r=None, w=None):
    """(first global env index, number of envs) of rank r when `global_envs` envs are split evenly."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    if global_envs % w != 0:
        raise ValueError(f"global env count {global_envs} must be divisible by the world size {w}")
    n = global_envs // w
    return r * n, n
