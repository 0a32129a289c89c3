"""Data-parallel plumbing of the hot path: mixtures are independent units,
so rank r takes rows [r*B/G, (r+1)*B/G) of every global batch (or its own synthetic stream), the
parameters are replicated, and the ONLY collective is one all-reduce(sum) of the flat fp32 gradient
buffer per step, scaled by 1/G inside the fused AMSGrad kernel.  Backend: NCCL over NVLink on the
GPUs; the same code runs over gloo on CPU tensors (tests/test_dp_gloo.py).  No torch.cuda use here."""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the launcher's environment (torchrun)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def bind_to_gpu_numa_node(pci_bus_id):
    """Pin this process (and with it the pages of the pinned host buffers it allocates next) to the CPUs of the NUMA node the
    GPU hangs off (`/sys/bus/pci/devices/<id>/numa_node`): with one process per GPU the H2D / D2H traffic of eight ranks
    otherwise crosses the socket interconnect at random.  Returns the node, or None when it cannot be determined
    (AMSS_NO_NUMA_BIND=1 switches it off).  Host-side plumbing only."""
    if os.environ.get("AMSS_NO_NUMA_BIND") == "1" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        dev = pci_bus_id.lower()
        if len(dev.split(":")[0]) == 8:                 # nvml style 00000000:1B:00.0 -> sysfs 0000:1b:00.0
            dev = dev[4:]
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError):
        return None


def active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_batch(batch, rank, world_size):
    """Rows [r*B/G, (r+1)*B/G) of every array of a global (mix, non_mix, ind) batch."""
    B = batch[0].shape[0]
    if B % world_size:
        raise ValueError(f"global batch {B} is not divisible by the world size {world_size}")
    per = B // world_size
    return tuple(a[rank * per:(rank + 1) * per] for a in batch)


def allreduce_sum_(flat):
    """The single collective of the path: in-place sum of the flat gradient buffer over all ranks.
    Returns the scale (1/G) the optimizer must apply."""
    if not active():
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()


class GradBuckets:
    """The gradient all-reduce of the path, overlapped with the backward pass: the flat gradient buffer is cut into
    contiguous buckets (one per layer, in the order the backward pass finishes them: head first, first BLSTM layer last),
    and a bucket's all-reduce(sum) is launched asynchronously the moment the last of its parameters has received its
    gradient (post-accumulate-grad hooks call ready()).  finish() launches whatever is left (parameters the loss did not
    reach keep a zero gradient), waits for every bucket and returns the 1/G scale for the optimizer.  Still ONE logical
    exchange of the gradient per step over the same bytes; on the GPUs the NCCL kernels are captured into the step's CUDA
    graph together with forward + backward.  Works on any backend (tests/test_dp_gloo.py drives it over gloo)."""

    def __init__(self, flat, buckets):
        """flat: the 1-D gradient buffer; buckets: [(lo, hi, n_params)] disjoint slices of it."""
        self.flat = flat
        self.buckets = [(int(lo), int(hi), int(n)) for lo, hi, n in buckets]
        self.reset()

    def reset(self):
        self.left = [n for _, _, n in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.works = []

    def _launch(self, i):
        lo, hi, _ = self.buckets[i]
        self.launched[i] = True
        if active() and hi > lo:
            self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def ready(self, i):
        """One more parameter of bucket i has its gradient."""
        self.left[i] -= 1
        if self.left[i] == 0 and not self.launched[i]:
            self._launch(i)

    def finish(self):
        for i in range(len(self.buckets)):
            if not self.launched[i]:
                self._launch(i)
        for w in self.works:
            w.wait()
        self.works = []
        return 1.0 / dist.get_world_size() if active() else 1.0


def clip_factor(norm_of_sum, clip, grad_scale=1.0):
    """Host statement of what amss_clip_factor computes on the device: tf.clip_by_global_norm (models/network.py:191-192)
    acts on the gradient of the batch-MEAN loss, while after the all-reduce the buffer holds the SUM over the G ranks,
    so the norm that is compared with `clip` is |grad_scale| * ||sum|| with grad_scale = 1/G.  The optimizer then
    multiplies the buffer by grad_scale * clip_factor."""
    return clip / max(abs(grad_scale) * norm_of_sum, clip)


def max_over_ranks(value, device="cpu"):
    """Timing helper: max of a python float over ranks (multi-GPU numbers are max-over-ranks)."""
    if not active():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
