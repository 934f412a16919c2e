"""Multi-GPU sharding of the stream batch: one process per GPU, streams partitioned by index.

The path has no cross-stream dependency (every stream owns its filter history and PLL state,
reference src/project.cpp:240-255), so there is NO data-path collective: rank g of G runs its
own Pipeline on streams [g*N/G, (g+1)*N/G).  torch.distributed is used only for the barrier,
the max-over-ranks timing and an optional host-side gather of the int16 audio to rank 0."""
import os


def stream_range(n_streams, world_size, rank):
    """Contiguous, balanced slice of the batch owned by `rank` (first n_streams % world_size ranks get one extra)."""
    assert 0 <= rank < world_size and n_streams >= 0
    base, extra = divmod(n_streams, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched directly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    next (first touch) are local to that GPU's PCIe root: with 8 ranks uploading at once a remote-socket staging
    buffer costs upload bandwidth.  Best effort: returns the node or None, never raises."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1 and torch.cuda.is_available():
        bind_to_gpu_numa_node(local_rank)
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class SharedRows:
    """The host-side gather of the path (the reference's sink is fwrite at src/project.cpp:317): ONE array
    [n_streams, row_len] in POSIX shared memory, mapped by every rank and registered with CUDA as pinned memory, so each rank's
    device->host copies land directly in its slice and rank 0 holds the whole batch's output after a barrier — no pickling, no
    second copy, no collective.  `mine` is this rank's slice (rows lo..hi of the job), `all` the whole array."""

    def __init__(self, name, n_streams, row_len, dtype, rank, world, pin=True):
        import numpy as np
        self.path = "/dev/shm/" + name
        self.rank, self.world = rank, world
        nbytes = int(n_streams) * int(row_len) * np.dtype(dtype).itemsize
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(max(nbytes, 1))
        barrier()
        self.all = np.memmap(self.path, dtype=dtype, mode="r+", shape=(int(n_streams), int(row_len)))
        lo, hi = stream_range(n_streams, world, rank)
        self.mine = self.all[lo:hi]
        self._registered = False
        if pin:
            try:
                import torch
                if torch.cuda.is_available() and nbytes:
                    rc = torch.cuda.cudart().cudaHostRegister(self.all.ctypes.data, nbytes, 0)
                    self._registered = int(rc) == 0
            except Exception:
                self._registered = False
        barrier()

    @property
    def pinned(self):
        return self._registered

    def close(self):
        barrier()
        if self._registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.all.ctypes.data)
            self._registered = False
        del self.mine
        del self.all
        barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        barrier()


def gather_rows_to_rank0(rows, n_streams):
    """Host-side gather of per-rank output rows (CPU tensor [n_local, L]) into [n_streams, L] on rank 0 through
    torch.distributed (pickled objects): a convenience for small results.  The timed path uses SharedRows."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return rows
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(rows, parts, dst=0)
    if rank != 0:
        return None
    out = torch.cat(parts, dim=0)
    assert out.shape[0] == n_streams
    return out
