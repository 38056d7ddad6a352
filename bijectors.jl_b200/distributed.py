"""Multi-GPU: columns are i.i.d. samples, so the batch shards by column with NO data-path collective;
the only exchange is one sum of the scalar batch log-density.  One process per GPU
(torchrun); torch.distributed is used for rendezvous only, the all-reduce itself is libb2b's NCCL call
site (b2b_allreduce_sum_f64).  On a CPU/gloo world (unit tests) the scalar is reduced with gloo."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import check, lib


def shard_columns(N: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous column block [lo, hi) of rank `rank`; blocks differ by at most one column."""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def numa_bind(device: Optional[int] = None) -> Tuple[int, int]:
    """Bind the calling thread (CPU affinity + preferred memory node) to the NUMA node of CUDA device `device`
    (default: the current device) BEFORE allocating pinned host batches: a pinned buffer on the other socket sends every
    H2D / D2H byte across the inter-socket link (b2b_numa_bind_to_device, include/b2b.h).  Returns (node, cpus); node is
    -1 when the topology is not exposed.  One process per GPU calls this once."""
    if device is None:
        device = torch.cuda.current_device()
    node, ncpu = ctypes.c_int32(-1), ctypes.c_int32(0)
    check(lib().b2b_numa_bind_to_device(int(device), ctypes.byref(node), ctypes.byref(ncpu)), "b2b_numa_bind_to_device")
    return int(node.value), int(ncpu.value)


def device_for_rank(local_rank: int, n_visible: Optional[int] = None) -> int:
    """CUDA device of local rank `local_rank`: the visible devices ordered round-robin over the NUMA nodes of their PCIe
    root complexes, so that a job with fewer ranks than GPUs spreads over the sockets (the host-buffer path is bound by
    host DRAM bandwidth per socket).  With every GPU in use, or without topology information, this is the identity."""
    n = torch.cuda.device_count() if n_visible is None else n_visible
    nodes = []
    for d in range(n):
        node = ctypes.c_int32(-1)
        try:
            check(lib().b2b_device_numa_node(d, ctypes.byref(node)), "b2b_device_numa_node")
        except Exception:
            node = ctypes.c_int32(-1)
        nodes.append(int(node.value))
    if any(v < 0 for v in nodes) or len(set(nodes)) <= 1:
        return local_rank % max(n, 1)
    per_node = {}
    for d, v in enumerate(nodes):
        per_node.setdefault(v, []).append(d)
    order, keys = [], sorted(per_node)
    while any(per_node[k] for k in keys):
        for k in keys:
            if per_node[k]:
                order.append(per_node[k].pop(0))
    return order[local_rank % n]


class Communicator:
    """b2b_comm wrapper: the NCCL unique id is created on rank 0 and broadcast through torch.distributed."""

    def __init__(self):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.handle = None
        if dist.get_backend() == "nccl" and torch.cuda.is_available():
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = ctypes.create_string_buffer(128)
                check(lib().b2b_comm_unique_id(buf), "b2b_comm_unique_id")
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            uid = uid.cuda()
            dist.broadcast(uid, src=0)
            raw = bytes(uid.cpu().tolist())
            h = ctypes.c_void_p()
            check(lib().b2b_comm_init_rank(ctypes.byref(h), self.world, self.rank, raw), "b2b_comm_init_rank")
            self.handle = h

    def allreduce_sum_(self, value: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks of a float64 tensor (device scalar on NCCL worlds)."""
        if value.dtype != torch.float64:
            raise TypeError("allreduce_sum_ expects float64")
        if self.handle is not None and value.is_cuda:
            check(lib().b2b_allreduce_sum_f64(self.handle, value.data_ptr(), value.numel(),
                                              torch.cuda.current_stream().cuda_stream), "b2b_allreduce_sum_f64")
        else:
            dist.all_reduce(value, op=dist.ReduceOp.SUM)
        return value

    def close(self):
        if self.handle is not None:
            lib().b2b_comm_destroy(self.handle)
            self.handle = None


def sharded_logpdf_sum(td, y_local: torch.Tensor, comm: Optional[Communicator]) -> torch.Tensor:
    """Σ over ALL ranks of logpdf(td, y) given this rank's column shard: local fused inverse-chain +
    MvNormal + reduction, then one all-reduce of 8 bytes."""
    from .transformed_distribution import logpdf_sum

    total, _ = logpdf_sum(td, y_local)
    if comm is not None and comm.world > 1:
        comm.allreduce_sum_(total.reshape(1))
    return total


def pack_param_grads(grads) -> torch.Tensor:
    """[{"w", "u", "b"}, ...] -> one float64 vector (w_1 | u_1 | b_1 | w_2 | ...): the payload of the training path's
    single exchange, the sum of the parameter cotangents over the column shards."""
    return torch.cat([torch.cat([g["w"].reshape(-1), g["u"].reshape(-1), g["b"].reshape(-1)]) for g in grads]).to(torch.float64)


def unpack_param_grads(buf: torch.Tensor, like) -> list:
    """Inverse of pack_param_grads (shapes / dtypes taken from `like`)."""
    out, off = [], 0
    for g in like:
        item = {}
        for k in ("w", "u", "b"):
            n = g[k].numel()
            item[k] = buf[off:off + n].to(g[k].dtype).reshape(g[k].shape)
            off += n
        out.append(item)
    return out


def sharded_planar_chain_vjp(t, x_local: torch.Tensor, ybar_local: torch.Tensor, ljbar_local, comm: Optional[Communicator]):
    """Reverse mode of a planar chain over column shards: every rank runs b2b_planar_chain_vjp_f32 on its columns
    (x̄ stays local -- it is a per-column quantity), then ONE all-reduce sums the 2·L·D + L parameter cotangents
    (data-parallel training's only exchange; the reference is single-process and has no counterpart)."""
    from .interface import planar_chain_vjp

    xbar, grads = planar_chain_vjp(t, x_local, ybar_local, ljbar_local)
    if comm is not None and comm.world > 1:
        buf = pack_param_grads(grads)
        comm.allreduce_sum_(buf)
        grads = unpack_param_grads(buf, grads)
    return xbar, grads


class Clique:
    """ONE process driving several GPUs (b2b_comm_init_all): per-device float64 scalars are summed across the devices of
    the calling process inside one NCCL group -- what a single Julia session holding all GPUs of a box would use."""

    def __init__(self, devices=None):
        n = torch.cuda.device_count() if devices is None else len(devices)
        self.devices = list(range(n)) if devices is None else [int(d) for d in devices]
        arr = (ctypes.c_int * n)(*self.devices)
        h = ctypes.c_void_p()
        check(lib().b2b_comm_init_all(ctypes.byref(h), n, arr), "b2b_comm_init_all")
        self.handle = h

    def allreduce_sum_(self, values):
        """values[i]: float64 CUDA tensor on devices[i] (same numel); reduced in place on every device."""
        n = len(self.devices)
        ptrs = (ctypes.c_void_p * n)(*[v.data_ptr() for v in values])
        streams = (ctypes.c_void_p * n)(*[torch.cuda.current_stream(d).cuda_stream for d in self.devices])
        check(lib().b2b_allreduce_sum_f64_all(self.handle, ptrs, values[0].numel(), streams), "b2b_allreduce_sum_f64_all")
        return values

    def close(self):
        if self.handle is not None:
            lib().b2b_comm_destroy(self.handle)
            self.handle = None
