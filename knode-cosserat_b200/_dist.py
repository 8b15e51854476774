"""Multi-GPU plumbing of the training path: one process per GPU, trajectories sharded across ranks, ONE collective per
optimiser step — an all-reduce(sum) of the flat 27,673-element MLP gradient.  Rollouts shard the rod
batch and need no collective at all.  Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests); no kernels here.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(device_index=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* when launched under torchrun.
    Returns (rank, world).  A plain `python script.py` run is rank 0 of 1 and initialises nothing."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            if device_index is None:
                device_index = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(device_index)
            bind_to_gpu_numa_node(device_index)
            dist.init_process_group("nccl", device_id=torch.device("cuda", device_index))
        else:
            dist.init_process_group("gloo")
    return rank, world


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and therefore its pinned host buffers, first touch) to the CPUs of the NUMA node the GPU hangs
    off.  With one rank per GPU and 410 MB of trajectory per step coming back over PCIe, ranks that all sit on one socket
    funnel every device-to-host stream through that socket's memory.  Best effort: returns the node or None."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world):
    """Contiguous, balanced partition of n items: rank r owns [lo, hi); sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(tensors, group=None):
    """In-place sum of a list of tensors across ranks through ONE flat buffer (one collective launch).  The loss
    normalisation 1/(batch_len-1) is a global constant (physics_train.py:267), so a plain sum reproduces the
    single-process loss and gradient."""
    rank, world = world_info()
    if world == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return tensors
