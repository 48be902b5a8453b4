"""Data-parallel training step for the cosmology benchmark on B200s.

Mirrors ``train_step`` of benchmarks/galaxies/train_cosmology.py:219-255 (reference):
build the kNN graph of the local batch, forward, MSE loss (:216,236-243), gradients,
mean of the gradients over devices (``jax.lax.pmean``, :246) and an ``optax.adamw`` update
with a cosine-decay learning rate (:439-440).  One process per GPU; the only collective is
one NCCL all-reduce (AVG) of the flat fp32 gradient bucket, with the loss riding in its
last element (the metric ``pmean`` of :253).  The reference's second, metrics-only forward
(:249-252) recomputes the same loss value and is not repeated.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch
import torch.distributed as dist

from . import graph_utils as GU
from . import ops
from .egnn import EGNN
from .equivariant_graph_utils import get_equivariant_graph
from .irreps import Irreps
from .nequip import IrrepsArray, NequIP, NequIPParams
from .segnn import SEGNN


def cosine_decay_lr(step: int, init_value: float, decay_steps: int, alpha: float = 0.0) -> float:
    """optax.cosine_decay_schedule."""
    c = min(step, decay_steps) / float(decay_steps)
    return init_value * ((1.0 - alpha) * 0.5 * (1.0 + math.cos(math.pi * c)) + alpha)


@dataclasses.dataclass
class TrainState:
    model: object                        # NequIP, EGNN or SEGNN (GraphWrapper of train_cosmology.py:136-212)
    params: NequIPParams
    learning_rate: float = 3e-4          # train_cosmology.py:645
    weight_decay: float = 1e-5           # :646
    decay_steps: int = 2000              # :439
    step: int = 0

    def __post_init__(self):
        n = self.params.flat.numel()
        dev = self.params.flat.device
        # gradient bucket with one extra slot for the loss: a single all-reduce per step
        self.bucket = torch.zeros((n + 1,), dtype=torch.float32, device=dev)
        self.params.grad = self.bucket[:n]
        self.mu = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.nu = torch.zeros((n,), dtype=torch.float32, device=dev)
        self._gpred: Optional[torch.Tensor] = None


def shard_batch(x, rank: int, world_size: int):
    """``split_batches`` of train_cosmology.py:283-298: rank r takes the r-th contiguous slice of the batch."""
    if x.shape[0] % world_size != 0:
        raise ValueError(f"batch of {x.shape[0]} graphs does not split over {world_size} devices")
    per = x.shape[0] // world_size
    return x[rank * per:(rank + 1) * per]


def allreduce_mean_(bucket: torch.Tensor, world_size: int) -> torch.Tensor:
    """``jax.lax.pmean(grads, "batch")`` (train_cosmology.py:246,253) on the flat gradient+loss bucket.
    NCCL averages in the collective; backends without AVG (gloo) sum and scale."""
    if world_size > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(bucket, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
            bucket.mul_(1.0 / world_size)
    return bucket


def wrap_nodes(halos: torch.Tensor) -> IrrepsArray:
    """GraphWrapper of train_cosmology.py:192-199: positions-only clouds are tiled to 1o+1o+1x0e."""
    if halos.shape[-1] == 3:
        ones = torch.ones(halos.shape[:2] + (1,), dtype=halos.dtype, device=halos.device)
        halos = torch.cat([halos, halos, ones], dim=-1)
    return IrrepsArray("1o + 1o + 1x0e", halos.contiguous())


def wrap_graph(model, halo_batch: torch.Tensor, senders, receivers, k: int, apply_pbc=None, globals_=None,
               n_radial_basis: int = 64, r_max: float = 0.6):
    """``GraphWrapper.__call__`` of train_cosmology.py:143-212: the model-specific view of one batch of clouds."""
    B, N = halo_batch.shape[:2]
    n_node = torch.full((B, 1), N, dtype=torch.int32)
    n_edge = torch.full((B, 1), k, dtype=torch.int32)
    if isinstance(model, NequIP):                            # :192-208
        return GU.GraphsTuple(nodes=wrap_nodes(halo_batch), edges=None, receivers=receivers, senders=senders,
                              globals=globals_, n_node=n_node, n_edge=n_edge)
    if isinstance(model, EGNN):                              # :150-151 (the model carries apply_pbc, :403-408)
        return GU.GraphsTuple(nodes=halo_batch.contiguous(), edges=None, receivers=receivers, senders=senders,
                              globals=globals_, n_node=n_node, n_edge=n_edge)
    if isinstance(model, SEGNN):                             # :163-190
        nodes = IrrepsArray(Irreps("1o" if halo_batch.shape[-1] == 3 else "1o + 1o"), halo_batch.contiguous())
        return get_equivariant_graph(nodes, halo_batch[..., :3].contiguous(), None, senders, receivers, n_node, n_edge, None,
                                     globals_, model.l_max_hidden, steerable_velocities=False, apply_pbc=apply_pbc,
                                     n_radial_basis=n_radial_basis, r_max=r_max)
    raise TypeError(f"no B200 path for model type {type(model).__name__}")


def local_gradients(state: TrainState, halo_batch: torch.Tensor, y_batch: torch.Tensor, k: int,
                    apply_pbc: Optional[GU.ApplyPBC] = None, globals_=None) -> torch.Tensor:
    """The rank-local part of a step: kNN graph, forward, MSE, backward.  Gradients and the loss land in state.bucket.
    No host synchronisation, no host-dependent launch parameter: capturable in a CUDA graph (GraphedTrainStep)."""
    model, params = state.model, state.params
    n = params.flat.numel()
    # build_graph (train_cosmology.py:226-235) also expands |dr| in a Bessel basis into graph.edges, which the
    # model wrappers discard (edges=None, :201-208); under jax.jit XLA removes that dead computation, so only
    # the neighbour search is run here.
    senders, receivers, _ = ops.knn_pbc(halo_batch, k, None if apply_pbc is None else apply_pbc.cell,
                                        None if apply_pbc is None else apply_pbc.cell_inv, want_dr=False)
    g = wrap_graph(model, halo_batch, senders, receivers, k, apply_pbc, globals_)
    loss_slot = state.bucket[n:]

    def grad_fn(out: torch.Tensor) -> torch.Tensor:
        if state._gpred is None or state._gpred.shape != out.shape:
            state._gpred = torch.empty_like(out)
        ops.mse_loss(out, y_batch, 1.0, loss_slot, state._gpred)
        return state._gpred

    model.forward_backward(params, g, grad_fn)
    return loss_slot


def apply_update(state: TrainState, world_size: int = 1) -> torch.Tensor:
    """pmean of gradients + loss (one NCCL all-reduce of the flat bucket), then optax.adamw with cosine decay."""
    params = state.params
    n = params.flat.numel()
    loss_slot = state.bucket[n:]
    allreduce_mean_(state.bucket, world_size)
    state.step += 1
    lr = cosine_decay_lr(state.step - 1, state.learning_rate, state.decay_steps)
    ops.adamw(params.flat, params.grad, state.mu, state.nu, lr, 0.9, 0.999, 1e-8, state.weight_decay, state.step)
    return loss_slot


def train_step(state: TrainState, halo_batch: torch.Tensor, y_batch: torch.Tensor, k: int,
               apply_pbc: Optional[GU.ApplyPBC] = None, world_size: int = 1, globals_=None) -> torch.Tensor:
    """One optimisation step on this rank's shard; returns the (device) mean loss over all ranks."""
    local_gradients(state, halo_batch, y_batch, k, apply_pbc, globals_)
    return apply_update(state, world_size)


class GraphedTrainStep:
    """The same step with the rank-local part (about 130 kernel launches for NequIP cfg-2) captured once in a CUDA
    graph and replayed: one graph launch + the all-reduce + the AdamW launch per step.  What the reference gets from
    jax.jit (train_cosmology.py:219: the whole step is one XLA executable).  Inputs are copied into static buffers
    (device-to-device or from pinned host memory); shapes are fixed at construction, like a jitted function's."""

    def __init__(self, state: TrainState, halo_shape, y_shape, k: int, apply_pbc: Optional[GU.ApplyPBC] = None,
                 world_size: int = 1, warmup: int = 2):
        self.state, self.k, self.apply_pbc, self.world_size = state, k, apply_pbc, world_size
        dev = state.params.flat.device
        self.halos = torch.zeros(tuple(halo_shape), dtype=torch.float32, device=dev)
        self.y = torch.zeros(tuple(y_shape), dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._warmup = warmup

    def capture(self, halo_batch: torch.Tensor, y_batch: torch.Tensor) -> None:
        """Warm up eagerly on a side stream (workspaces reach their final size), then capture."""
        self.halos.copy_(halo_batch)
        self.y.copy_(y_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                local_gradients(self.state, self.halos, self.y, self.k, self.apply_pbc)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            local_gradients(self.state, self.halos, self.y, self.k, self.apply_pbc)
        self.graph = g

    def __call__(self, halo_batch: torch.Tensor, y_batch: torch.Tensor) -> torch.Tensor:
        if self.graph is None:
            self.capture(halo_batch, y_batch)
        self.halos.copy_(halo_batch, non_blocking=True)
        self.y.copy_(y_batch, non_blocking=True)
        self.graph.replay()
        return apply_update(self.state, self.world_size)
