"""Query sharding across the GPUs of one node: one process per GPU, every rank holds the whole packed
reference block and works on a contiguous, letter-balanced range of the query block.  The ONLY collective is the one-off
broadcast of the packed reference block (NCCL over NVLink on GPUs, gloo in the CPU tests); there is no collective on the
data path and results are concatenated in rank order."""
from __future__ import annotations
import numpy as np


def pin_to_gpu_numa_node(gpu_index: int, pci_bus_id: str | None = None) -> str:
    """Restricts this process (and the host threads the library starts later) to the CPUs of the NUMA node the GPU hangs on: the
    host side of a rank -- pinned staging buffers, the worker pool that scores and culls -- then stays next to its GPU's PCIe root.
    On a two-socket 8-GPU box GPUs 4-7 sit on node 1; unpinned ranks of those GPUs ran their host work across the socket link.
    Returns a short note for the bench record; does nothing (and says why) where the topology cannot be read."""
    import os, subprocess
    try:
        # pci_bus_id ("0000:1b:00.0") of the CUDA device if the caller knows it (CUDA's device order need not be nvidia-smi's); else, or if
        # sysfs does not know that address, ask nvidia-smi for the device of this index
        node = None
        for src in ("caller", "nvidia-smi"):
            bus = pci_bus_id if src == "caller" else subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                                                     capture_output=True, text=True, timeout=20).stdout.strip()
            if not bus:
                continue
            dom, rest = bus.split(":", 1)
            path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
            if os.path.exists(path):
                node = int(open(path).read())
                break
        if node is None:
            return "no NUMA information for the device"
        if node < 0:
            return "single NUMA node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node}: no allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # topology files missing in a container, nvidia-smi absent, ...
        return f"not pinned ({type(e).__name__})"


def query_ranges(q_limits: np.ndarray, parts: int):
    """Contiguous query-id ranges with ~equal letter counts (the reference's SequenceSet::partition,
    data/sequence_set.cpp:57-75, does the same for threads)."""
    n = len(q_limits) - 1
    total = int(q_limits[-1] - q_limits[0])
    cuts = [0]
    for k in range(1, parts):
        want = int(q_limits[0]) + total * k // parts
        c = int(np.searchsorted(q_limits, want, side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(parts)]


def sub_block(raw: np.ndarray, limits: np.ndarray, begin: int, end: int):
    """Block image of sequences [begin, end) cut out of a block image (ids are re-based to 0)."""
    from .api import PADDING, DELIMITER
    a, b = int(limits[begin]), int(limits[end])
    out = np.full(b - a + 2 * PADDING, DELIMITER, dtype=np.int8)
    out[PADDING:PADDING + (b - a)] = raw[a:b]
    lim = (limits[begin:end + 1] - a + PADDING).astype(np.int64)
    return out, lim


def broadcast_reference(r_raw, r_limits, dist, device=None):
    """Rank 0's packed reference block to every rank (one broadcast of the letters, one of the limits)."""
    import torch
    rank = dist.get_rank()
    meta = torch.tensor([len(r_raw) if rank == 0 else 0, len(r_limits) if rank == 0 else 0], dtype=torch.int64)
    if device is not None:
        meta = meta.to(device)
    dist.broadcast(meta, 0)
    n_raw, n_lim = int(meta[0]), int(meta[1])
    t = torch.from_numpy(np.ascontiguousarray(r_raw).view(np.uint8)) if rank == 0 else torch.empty(n_raw, dtype=torch.uint8)
    l = torch.from_numpy(np.ascontiguousarray(r_limits)) if rank == 0 else torch.empty(n_lim, dtype=torch.int64)
    if device is not None:
        t, l = t.to(device), l.to(device)
    dist.broadcast(t, 0)
    dist.broadcast(l, 0)
    return t.cpu().numpy().view(np.int8), l.cpu().numpy()


def broadcast_reference_block(ctx, dist, device, block=None, root: int = 0):
    """The resident reference block of rank `root` on every rank, sent device to device by the library's own NCCL broadcast
    (dmnd_block_broadcast, csrc/cuda/comm.cu): masking, soft table and bias travel with the letters, nothing bounces through the
    host.  torch.distributed only carries the 128-byte NCCL unique id.  Returns (block, raw_len, limits)."""
    import os, torch
    if "DMND_NCCL_LIB" not in os.environ:  # the NCCL torch was built with (the system one may be older)
        try:
            import nvidia.nccl
            cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["DMND_NCCL_LIB"] = cand
        except Exception:
            pass
    rank = dist.get_rank()
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == root:
        uid = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(uid, root)
    ctx.comm_init(rank, dist.get_world_size(), bytes(uid.cpu().numpy().tobytes()))
    return ctx.block_broadcast(root, block if rank == root else None)


def gather_matches(matches: np.ndarray, query_offset: int, dist):
    """All ranks' match records on rank 0, query ids re-based to the unsharded block, in rank (= query) order."""
    m = matches.copy()
    m["query"] += query_offset
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(m, out, dst=0)
    return np.concatenate(out) if dist.get_rank() == 0 else None
