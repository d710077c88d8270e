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


class LayerwiseReducer:
    """Gradient all-reduce that starts while the backward pass is still running: as soon as a layer's slice of the flat
    gradient arena is final (last backward pass over that network, the layer's weight-gradient second stage done) it is
    SUM-all-reduced on a communication stream, next to the remaining layers' kernels.  The PatchGAN's 512 -> 1024 conv
    (three quarters of the 44 MB arena) is the first layer of the backward pass, so its transfer hides behind the other
    four layers.  The stream forks from / joins the issuing stream with events, i.e. it is captured into the step's CUDA
    graph like the other side streams (NCCL collectives are capturable)."""

    def __init__(self, world_size):
        self.world_size, self.stream, self.used = world_size, None, False

    def reduce_async(self, t):
        if self.world_size <= 1:
            return
        self.used = True
        if not t.is_cuda:                       # gloo / CPU tests: synchronous
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return
        import torch
        cur = torch.cuda.current_stream()
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=t.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def join(self):
        if self.stream is not None and self.used:
            import torch
            torch.cuda.current_stream().wait_stream(self.stream)
