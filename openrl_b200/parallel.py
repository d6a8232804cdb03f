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


# ---- low-latency all-reduce of the small gradient bucket over NVLink peer memory ---------------------------------
# The bucket is 37 KB, i.e. latency-bound.  Optional path (ORL_SYMM_ALLREDUCE=1): torch's symmetric-memory one-shot
# all-reduce (every rank reads all peers' buffers over NVLink / NVSwitch and sums locally, signal-pad barriers,
# graph-capturable) instead of NCCL.  Measured on 2 x B200 inside the captured iteration graph: 1.862 ms vs 1.866 ms per
# iteration — no gain once the launches are graph-replayed, so NCCL (the north-star's "single NCCL allreduce on the
# gradient bucket per update") stays the default.  Any failure to set the symmetric path up falls back to NCCL.
_SYMM = {}   # data_ptr -> (symmetric input tensor, output tensor, group name)


def symmetric_buffer(shape, dtype, device):
    """A tensor to be SUM-all-reduced with `allreduce_sum_into`: symmetric memory when available (> 1 rank, float32),
    else a plain tensor.  Returns (buffer, reduced) — `reduced` receives the sum (the buffer itself in the NCCL path)."""
    import os

    import torch

    t = None
    if is_distributed() and dtype == torch.float32 and os.environ.get("ORL_SYMM_ALLREDUCE", "0") == "1":
        try:
            import torch.distributed._symmetric_memory as symm_mem

            group = dist.group.WORLD
            t = symm_mem.empty(*shape, dtype=dtype, device=device)
            t.zero_()
            symm_mem.rendezvous(t, group.group_name)
            out = torch.zeros(*shape, dtype=dtype, device=device)
            torch.ops.symm_mem.one_shot_all_reduce_out(t, "sum", group.group_name, out)   # first use: sets up / validates the path
            torch.cuda.synchronize()
            _SYMM[t.data_ptr()] = (t, out, group.group_name)
            return t, out
        except Exception as e:  # noqa: BLE001
            import warnings

            warnings.warn(f"openrl_b200: symmetric-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
            t = None
    import torch as _t

    t = _t.zeros(*shape, dtype=dtype, device=device)
    return t, t


def allreduce_sum_into(t):
    """SUM all-reduce of a buffer from `symmetric_buffer`; returns the tensor that holds the result."""
    ent = _SYMM.get(t.data_ptr())
    if ent is not None and is_distributed():
        import torch

        torch.ops.symm_mem.one_shot_all_reduce_out(ent[0], "sum", ent[2], ent[1])
        return ent[1]
    return allreduce_sum_(t)


def env_shard(global_envs, r=None, w=None):
    """(first global env index, number of envs) of rank r when `global_envs` envs are split evenly."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    if global_envs % w != 0:
        raise ValueError(f"global env count {global_envs} must be divisible by the world size {w}")
    n = global_envs // w
    return r * n, n
