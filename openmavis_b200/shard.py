"""Multi-GPU host logic: one process per GPU, rig frames sharded, no data-path collective for extraction; the cross-GPU
brute-force 2-NN is query-sharded with ONE all-gather of the train descriptor rows.

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU test-suite (where the 2-NN
compute callback is the oracle — the product callback is ORBmatcher.knnMatch2_device)."""
import numpy as np


def frame_shard(n_frames, rank, world):
    """Contiguous block of rig frames owned by `rank` (all cameras of a rig frame stay on one GPU, so the
    left/right 2-NN that follows extraction is local).  Returns (first, count); blocks differ by at most one frame."""
    base, rem = divmod(int(n_frames), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def row_shard(n_rows, rank, world):
    return frame_shard(n_rows, rank, world)


def gather_train_rows(local_train, n_total, dist, device=None):
    """all_gather of the train descriptor rows ([n_local, 32] uint8 torch tensor per rank, possibly ragged) -> [n_total, 32]."""
    import torch
    world = dist.get_world_size()
    counts = [row_shard(n_total, r, world)[1] for r in range(world)]
    assert local_train.shape[0] == counts[dist.get_rank()]
    m = max(counts)
    padded = local_train
    if local_train.shape[0] < m:                       # all_gather needs equal shapes: pad the short shards
        padded = torch.zeros((m, 32), dtype=torch.uint8, device=local_train.device)
        padded[:local_train.shape[0]] = local_train
    parts = [torch.empty((m, 32), dtype=torch.uint8, device=local_train.device) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def knn2_sharded(query_local, train_local, n_total, dist, knn2_fn):
    """Brute-force 2-NN of this rank's query rows against ALL train rows.  knn2_fn(query, train) -> (idx[nq,2], dist[nq,2]) with
    GLOBAL train indices, ordered (distance asc, trainIdx asc) like cv::BFMatcher.knnMatch (src/Frame.cc:1427)."""
    train = gather_train_rows(train_local, n_total, dist)
    return knn2_fn(query_local, train)


def gather_results(idx_local, dist_local, n_total, dist):
    """Optional gather of the sharded results to every rank (numpy in, numpy out)."""
    import torch
    world = dist.get_world_size()
    counts = [row_shard(n_total, r, world)[1] for r in range(world)]
    m = max(counts)
    buf = torch.full((m, 4), -1, dtype=torch.int32)
    n = idx_local.shape[0]
    buf[:n, :2] = torch.from_numpy(np.ascontiguousarray(idx_local, np.int32))
    buf[:n, 2:] = torch.from_numpy(np.ascontiguousarray(dist_local, np.int32))
    parts = [torch.empty((m, 4), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)]).numpy()
    return out[:, :2].copy(), out[:, 2:].copy()


def host_numa_nodes():
    """Number of NUMA nodes the host exposes (1 on a single-socket or virtualised box), or None when /sys cannot be read."""
    import os
    try:
        return len([d for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit()])
    except OSError:
        return None


def bind_to_gpu_numa(pci_bus_id):
    """Pin this process (one per GPU) to the CPU cores of the NUMA node its GPU hangs off, so that the pinned staging buffers it
    allocates afterwards are first-touched on that node and the host<->device copies do not cross the socket interconnect.
    pci_bus_id: 'dddd:bb:dd.f' (e.g. from torch.cuda.get_device_properties(i): '%04x:%02x:%02x.0' % (pci_domain_id, pci_bus_id,
    pci_device_id)).  Returns the node number, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        with open('/sys/bus/pci/devices/%s/numa_node' % pci_bus_id.lower()) as f:
            node = int(f.read().strip())
        if node < 0:
            # the device reports no node (virtual machines do): with a single node there is nothing to bind — report node 0;
            # otherwise fall back to the CPU list the PCI device names as local
            if host_numa_nodes() == 1:
                return 0
            with open('/sys/bus/pci/devices/%s/local_cpulist' % pci_bus_id.lower()) as f:
                spec = f.read().strip()
            cpus = set()
            for part in spec.split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if not allowed or allowed == os.sched_getaffinity(0):
                return None
            os.sched_setaffinity(0, allowed)
            return -1
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None
