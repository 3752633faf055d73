"""Multi-GPU plumbing of the registration sweep: pairs are independent, so rank r of W registers the
pairs r, r+W, ... with replicated weights and NO data-path collective; the per-pair metric rows are exchanged with ONE
all_gather at the end (the reference does one all_reduce per scalar: engine/base_trainer.py:229-234, utils/torch.py:16-21).
Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def pair_ids(n_local, rank, world):
    """global ids of the n_local pairs rank `rank` registers: rank + i * world"""
    return [rank + i * world for i in range(n_local)]


def gather_metric_rows(rows, world):
    """rows: (n_local, k) float tensor on this rank's device -> (world * n_local, k), rank-major, via one all_gather"""
    if world <= 1:
        return rows
    gathered = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(gathered, rows)
    return torch.cat(gathered, dim=0)


def max_over_ranks(value, device, world):
    """the timing reduction of the bench contract: max of a per-rank scalar over all ranks"""
    if world <= 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
