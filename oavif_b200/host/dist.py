"""One process per GPU, no collective on the data path (images are independent — SURVEY.md §8e).

torch.distributed is used only for the plumbing around the timed region: rendezvous, barrier, the
max-over-ranks of the elapsed time and the host-side merge of per-image rows.  NCCL on the GPU box,
gloo in the CPU tests (world_size 2)."""
from __future__ import annotations

import os


def env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str | None = None):
    """Join the job described by RANK/WORLD_SIZE/MASTER_* (torchrun).  Returns (rank, world, local)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def _parse_cpulist(text: str) -> list[int]:
    out: list[int] = []
    for tok in text.strip().split(","):
        if not tok:
            continue
        a, _, b = tok.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def gpu_local_cpus(pci_bus_id: str) -> list[int]:
    """CPUs the kernel reports as local to a PCI function (sysfs local_cpulist); [] when unknown."""
    try:
        with open(f"/sys/bus/pci/devices/{pci_bus_id.lower()}/local_cpulist") as f:
            return _parse_cpulist(f.read())
    except OSError:
        return []


def bind_near_gpu(local_rank: int, local_world: int, pci_bus_id: str | None = None) -> dict:
    """Pin this process (and the threads it starts later) BEFORE it allocates pinned host memory, so that the
    staging buffers are first-touched on the memory node next to the GPU's PCIe root and the caller threads
    stay there.  Uses the GPU's sysfs local_cpulist when it names a proper subset of the visible CPUs;
    otherwise (one NUMA node visible, as on the single-socket VMs of this pool) the visible CPUs are split
    evenly between the local ranks, which at least stops ranks from migrating across each other's caches.
    Returns what was done, for the bench line."""
    try:
        visible = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return {"policy": "unavailable"}
    local = [c for c in gpu_local_cpus(pci_bus_id) if c in visible] if pci_bus_id else []
    if local and len(local) < len(visible):
        share = [c for i, c in enumerate(local)]   # ranks that share a node share its CPUs
        os.sched_setaffinity(0, share)
        return {"policy": "gpu-local", "cpus": f"{share[0]}-{share[-1]}", "n": len(share)}
    if local_world > 1 and len(visible) >= local_world:
        per = len(visible) // local_world
        mine = visible[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return {"policy": "even-split", "cpus": f"{mine[0]}-{mine[-1]}", "n": len(mine)}
    return {"policy": "none", "n": len(visible)}


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Image i belongs to rank i mod world — the corpus driver's rule (oavif_host.cpp run_corpus)."""
    return list(range(rank, n_items, world))


def gather_rows(rows: list[tuple]) -> list[tuple] | None:
    """Per-image rows (index first) from every rank, merged in image order on rank 0."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(rows)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(rows, out, dst=0)
    if dist.get_rank() != 0:
        return None
    return sorted(r for part in out for r in part)
