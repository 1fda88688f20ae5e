"""Multi-GPU plumbing: one process per GPU, sample passes sharded across ranks, one sum-reduce of the
fp32 HDR accumulators at read-out.  torch.distributed is used for rendezvous and the
collective only; nothing here touches the light-transport kernels.

The path shards by ITERATION: iteration k depends only on (scene, camera, k) -- FrameSeed =
WangHash(k) (renderersystem.cpp:101), light samples re-seeded from k -- and iterations interact only
through `+=` into the accumulator (pathtracer.comp:107-109, screenreproject.glsl:45-48).  Rank r of W
renders k = r, r+W, r+2W, ...; the final image is sum(accumulators) / sum(iterations).
"""
import os

import numpy as np


def env_rank_world():
    """RANK / WORLD_SIZE / LOCAL_RANK as torchrun exports them (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_iterations(total_iterations, rank, world):
    """Global iteration indices rendered by `rank`: round-robin, disjoint across ranks, union = range(total)."""
    return list(range(rank, total_iterations, world))


def global_iteration(local_index, rank, world):
    """Global index of a rank's local_index-th iteration (RendererSystem::GetGlobalIterationIndex)."""
    return local_index * world + rank


def init_process_group(backend=None):
    """Join the job's process group (no-op for a single process).  MASTER_ADDR must be 127.0.0.1 on one node."""
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world == 1:
        return None
    if not dist.is_initialized():
        if backend is None:
            import torch
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            import torch
            # device_id: the communicator is created here (eagerly, on this rank's GPU) instead of inside the first collective
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", torch.cuda.current_device()))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def exchange_handles(handle, dist=None):
    """All ranks' 64-byte exchange-block handles in rank order (the transport gi_comm_connect leaves to the caller): one all_gather_object.
    Works on any backend (gloo on CPU for the tests, nccl on the GPU box -- object collectives go through the process group's store / CPU)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(handle)]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(handle))
    assert all(isinstance(h, bytes) and len(h) == 64 for h in out)
    return out


def connect_ranks(device_obj, dist=None):
    """gi_comm_create + handle exchange + gi_comm_connect for this rank's context; after it, Device.read_accumulator_reduced[_async] /
    RendererSystem.ReadBackbufferReduced[Async] are available.  Call after the screen size is set.  Returns (rank, world)."""
    rank, world, _ = env_rank_world()
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    handle = device_obj.comm_create(rank, world)
    device_obj.comm_connect(exchange_handles(handle, dist))
    if world > 1:
        dist.barrier()   # every rank has mapped every block before anyone issues the first reduced read
    return rank, world


def reduce_accumulator(accum_tensor, iteration_count, dist=None):
    """Sum-reduce the RGBA32F accumulators and the iteration counts over all ranks, in place.

    accum_tensor: torch tensor (CUDA for nccl, CPU for gloo) viewing the rank's accumulator.
    Returns the global iteration count.  One all_reduce of W*H*4 floats + one of a scalar.
    """
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(iteration_count)
    import torch
    dist.all_reduce(accum_tensor, op=dist.ReduceOp.SUM)
    n = torch.tensor([int(iteration_count)], dtype=torch.int64, device=accum_tensor.device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return int(n.item())


class DeviceBuffer:
    """Expose a raw device pointer (gi_accumulator_device) to torch through __cuda_array_interface__ so
    that NCCL can reduce the accumulator in place, without a copy."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def accumulator_as_tensor(device_obj, width, height):
    """torch view (no copy) of the context's accumulator: shape (H, W, 4) float32 on the context's GPU."""
    import torch
    ptr, nbytes = device_obj.accumulator_device()
    assert nbytes == width * height * 16
    return torch.as_tensor(DeviceBuffer(ptr, (height, width, 4)), device="cuda")


def rel_mse(image, reference):
    """relMSE = mean_px sum_rgb (x - r)^2 / (r^2 + 1e-2)   (SURVEY.md 8d)."""
    x = np.asarray(image, dtype=np.float64)[..., :3]
    r = np.asarray(reference, dtype=np.float64)[..., :3]
    return float(np.mean(np.sum((x - r) ** 2 / (r * r + 1e-2), axis=-1)))


def mse(image, reference):
    """Reference MSE definition: mean over pixels of sum_rgb (x - r)^2 (sqerror.frag:16-17, texturemse.cpp:58-62)."""
    x = np.asarray(image, dtype=np.float64)[..., :3]
    r = np.asarray(reference, dtype=np.float64)[..., :3]
    return float(np.mean(np.sum((x - r) ** 2, axis=-1)))
