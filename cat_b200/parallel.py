"""Data-parallel recipe of the distillation step (one process per GPU, torch.distributed).

The reference uses single-process ``nn.DataParallel`` (models/networks.py:160-161): L1 / GAN losses are
means over the *gathered* global batch, the KA term is computed per replica and **summed**
(distillers/inception_distiller.py:137-148), BatchNorm statistics stay per replica.  With equal shards the
same gradient is obtained by: local losses as usual, local KA scaled by ``world_size``, one SUM all-reduce
of each flat gradient arena, and a 1/world_size factor applied inside the Adam kernel.
"""
import torch.distributed as dist


def ka_scale(world_size: int) -> float:
    """Factor on the local KA term so that the mean-reduced gradient equals the reference's replica sum."""
    return float(world_size)


def grad_scale(world_size: int) -> float:
    """Factor applied to the SUM-all-reduced gradient arena (catb_adam's grad_scale argument)."""
    return 1.0 / float(world_size)


def reduce_gradients(flat_grad, world_size: int):
    """One collective per optimiser: SUM all-reduce of the flat fp32 gradient arena (NCCL on GPUs, gloo in
    the CPU tests).  Returns the tensor for chaining; a no-op for a single process."""
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad
