"""Multi-GPU plumbing: environments shard with no data-path collective (they never interact -- the reference
runs them in separate processes, baselines/baselines/common/vec_env/subproc_vec_env.py:50-54).  One process per
GPU owns the env-id range [env0, env0+n); seeds derive from the GLOBAL env id so a rollout does not depend on the
GPU count.  The only collective is the episode-statistics vector (the role of baselines' Monitor,
baselines/baselines/bench/monitor.py:58-76): 4 x int64, summed (max for the last entry) across ranks."""
import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (its PCIe root's NUMA node), so that
    pinned host buffers allocated afterwards are first-touched there: host-facing calls (tbx_step_host) then copy over the GPU's own
    root port instead of the inter-socket link.  Returns the CPU list it bound to, or None when NVML / affinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].strip().isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard(total_envs, rank, world):
    """Contiguous, balanced env-id range of `rank`: returns (env0, n)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(total_envs), int(world))
    n = base + (1 if rank < extra else 0)
    env0 = rank * base + min(rank, extra)
    return env0, n


def global_seeds(base_seed, env0, n):
    """Toybox.set_seed values of envs env0..env0+n-1: base_seed + global env id (u32 wrap)."""
    return ((int(base_seed) + env0 + np.arange(n, dtype=np.int64)) & 0xFFFFFFFF).astype(np.uint32)


def reduce_episode_stats(stats, device=None, group=None):
    """All-reduce [episodes, sum_return, sum_length, max_return] over the process group (NCCL on GPUs, gloo on CPU).
    Returns the global vector as python ints on every rank.  Without an initialised group it is the identity."""
    vec = [int(v) for v in stats]
    if len(vec) != 4:
        raise ValueError("episode statistics are [episodes, sum_return, sum_length, max_return]")
    if not (dist.is_available() and dist.is_initialized()):
        return vec
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sums = torch.tensor(vec[:3], dtype=torch.int64, device=device)
    mx = torch.tensor(vec[3:], dtype=torch.int64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return [int(sums[0]), int(sums[1]), int(sums[2]), int(mx[0])]
