"""Multi-GPU partitioning of the objective: one process per GPU, torch.distributed for the plumbing.

Two modes:

* candidates -- the batch of coefficient vectors is split into contiguous blocks, rank g evaluates
  block g over the whole k-grid.  Candidates are independent, so the only exchange is ONE
  all-gather of the (4 x B/G) result blocks (<= 128 KiB for B = 4096; latency-bound, NVLink
  bandwidth is irrelevant).  Replaces the reference's mp.Pool scan (optimize.py:169-190).
* slabs -- every rank evaluates ALL candidates over its own x-planes of the first (slowest) axis; no
  halo, the objective has no neighbour coupling.  ONE all-gather of the (4 x B) partials followed by
  a fixed rank-order combine (sum / min / max / sum) -- deterministic, NaN-propagating; on GPUs the
  combine is one tiny kernel of the library (sb200_combine_slabs_device), so a step is three
  launches: objective, all-gather, combine.  This is what makes a 2048^3 grid (64 GiB per NumPy
  array in the reference, i.e. not representable there) a 1/8-per-GPU job.  The planes are dealt
  round-robin (rank g owns planes g, g+G, g+2G, ...): the cost of a k-point depends on |k|, so
  contiguous slabs [gN/G, (g+1)N/G) -- mode 'slabs-contiguous', SURVEY.md 8e's wording -- leave the
  rank with the high-|k| slab working longest, and the step takes as long as the slowest rank.

The collective is torch.distributed's (NCCL on GPUs, gloo in the CPU tests).  The evaluator is
injected: on GPUs it is `Context.objective_batch_device`; the CPU tests (tests/test_sharding_gloo.py)
inject the oracle to exercise exactly this partition/combine logic at world_size 2.
"""
import numpy as np


def split_range(n, world, rank):
    """Contiguous block `rank` of range(n) split into `world` near-equal blocks (first blocks larger)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def block_sizes(n, world):
    return [split_range(n, world, r)[1] - split_range(n, world, r)[0] for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def combine_slab_partials(parts):
    """parts: tensor/array [world, 4, B] of per-rank (S, Amin, Amax, nan) -> [4, B], fixed rank order.

    S and nan are summed rank 0 first; Amin / Amax use min / max; a NaN in any rank's entry poisons
    that candidate's result, as numpy's reductions over the whole grid would."""
    import torch
    parts = torch.as_tensor(parts)
    out = parts[0].clone()
    for r in range(1, parts.shape[0]):
        out[0] = out[0] + parts[r, 0]
        out[3] = out[3] + parts[r, 3]
        nan = torch.isnan(out[1]) | torch.isnan(parts[r, 1]) | torch.isnan(out[2]) | torch.isnan(parts[r, 2])
        out[1] = torch.where(nan, torch.full_like(out[1], float("nan")), torch.minimum(out[1], parts[r, 1]))
        out[2] = torch.where(nan, torch.full_like(out[2], float("nan")), torch.maximum(out[2], parts[r, 2]))
    return out


class ShardedObjective:
    """Evaluate a batch of coefficient vectors across the ranks of a process group.

    evaluate_local(coeffs_block: torch.Tensor[b, ncoef], x_begin, x_end, x_stride) -> torch.Tensor[4, b]
    must return (S, Amin, Amax, nan-as-double) rows for its block over planes range(x_begin, x_end, x_stride)
    on the tensor's device.  combine(parts[world, 4, B]) -> [4, B] replaces the torch combine (GPU: one kernel).
    """

    MODES = ("candidates", "slabs", "slabs-contiguous")

    def __init__(self, evaluate_local, nx, mode="candidates", group=None, combine=None):
        if mode not in self.MODES:
            raise ValueError("mode must be one of %s" % (self.MODES,))
        self.evaluate_local = evaluate_local
        self.combine = combine or combine_slab_partials
        self.nx = int(nx)
        self.mode = mode
        self.group = group
        self.rank, self.world = world_info(group)

    def my_candidates(self, B):
        return split_range(B, self.world, self.rank) if self.mode == "candidates" else (0, B)

    def my_planes(self):
        """(x_begin, x_end, x_stride) of this rank."""
        if self.mode == "slabs":
            return self.rank, self.nx, self.world
        if self.mode == "slabs-contiguous":
            return split_range(self.nx, self.world, self.rank) + (1,)
        return 0, self.nx, 1

    def evaluate(self, coeffs):
        """coeffs: torch.Tensor [B, ncoef] (same on every rank) -> torch.Tensor [4, B] on every rank."""
        import torch
        B = coeffs.shape[0]
        lo, hi = self.my_candidates(B)
        x0, x1, xs = self.my_planes()
        if self.world > 1 and self.mode != "candidates" and x1 <= x0:
            raise ValueError("more ranks (%d) than x-planes (%d)" % (self.world, self.nx))
        if hi > lo:
            local = self.evaluate_local(coeffs[lo:hi].contiguous(), x0, x1, xs)
        else:
            local = torch.empty((4, 0), dtype=torch.float64, device=coeffs.device)
        if self.world == 1:
            return local
        dist = _dist()
        if self.mode == "candidates":
            sizes = block_sizes(B, self.world)
            bmax = max(sizes)
            if min(sizes) == bmax:
                padded = local.contiguous()
            else:
                padded = torch.zeros((4, bmax), dtype=torch.float64, device=coeffs.device)
                padded[:, :hi - lo] = local
            # output laid out as the concatenation along dim 0 (the form both NCCL and gloo accept)
            gathered = torch.empty((self.world * 4, bmax), dtype=torch.float64, device=coeffs.device)
            dist.all_gather_into_tensor(gathered, padded, group=self.group)
            gathered = gathered.view(self.world, 4, bmax)
            if min(sizes) == bmax:
                return gathered.permute(1, 0, 2).reshape(4, B)
            return torch.cat([gathered[r, :, :sizes[r]] for r in range(self.world)], dim=1)
        gathered = torch.empty((self.world * 4, B), dtype=torch.float64, device=coeffs.device)
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=self.group)
        return self.combine(gathered.view(self.world, 4, B))


def device_evaluator(ctx, scan=False, fast=False):
    """evaluate_local for a CUDA `Context`: device tensors in, device tensor out, on torch's stream.
    scan / fast: the kernel variant (start-grid screening; the shorter series of throughput launches)."""
    import torch

    def run(block, x_begin, x_end, x_stride=1):
        if not block.is_cuda or block.dtype != torch.float64:
            raise TypeError("device_evaluator needs float64 CUDA tensors")
        out = torch.empty((4, block.shape[0]), dtype=torch.float64, device=block.device)
        ctx.objective_batch_device(block.data_ptr(), block.shape[0], out.data_ptr(), x_begin, x_end,
                                   stream=torch.cuda.current_stream(block.device).cuda_stream, x_stride=x_stride,
                                   scan=scan, fast=fast)
        return out

    return run


def device_combiner(ctx):
    """combine for a CUDA `Context`: the library's fixed-order combine kernel on torch's stream."""
    import torch

    def run(parts):
        world, four, B = parts.shape
        out = torch.empty((4, B), dtype=torch.float64, device=parts.device)
        ctx.combine_slabs_device(parts.data_ptr(), world, B, out.data_ptr(),
                                 stream=torch.cuda.current_stream(parts.device).cuda_stream)
        return out

    return run


def host_evaluate(sharded, coeffs_host, device):
    """End-to-end call with HOST buffers: pinned H2D of this rank's inputs, kernels, the collective,
    D2H of the combined result.  Returns numpy (S, Amin, Amax, nan)."""
    import torch
    c = torch.from_numpy(np.ascontiguousarray(coeffs_host, dtype=np.float64))
    if device.type == "cuda":
        c = c.pin_memory().to(device, non_blocking=True)
    out = sharded.evaluate(c).cpu().numpy()
    return out[0], out[1], out[2], out[3].astype(np.int64)
