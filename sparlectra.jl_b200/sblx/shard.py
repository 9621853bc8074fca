"""Scenario sharding across ranks (one process per GPU) and the single collective of the path:
an all-gather of fixed-size per-scenario result records (64 bytes = SBLX_RECORD_BYTES: int32 converged, status,
iterations, n_switch; float64 final_mismatch, vmin, vmax, max_loading; int32 n_voltage_violations, n_overloaded,
island_count, n_tiny_pivots).

The PRODUCT path does this inside libsblx.so (sblx_solve_outages_sharded / sblx_multi_solve_outages: ncclAllGather
on the engine's stream).  This module is the host-side mirror of the same layout and bounds, used by the CPU
(gloo) tests and by callers that want to gather `sblx_device_summary` buffers themselves.

The reference fans contingencies out in contiguous chunks with no communication between workers
and returns results in input order (src/contingency.jl:299-318); ranks do the same with
`shard_range`, and `gather_records` restores the global input order.  Backend: NCCL on GPUs
(records stay in HBM: the tensor wraps the engine's device buffer), gloo in the CPU tests."""
from __future__ import annotations

import numpy as np

RECORD_BYTES = 64
RECORD_DTYPE = np.dtype([("converged", "<i4"), ("status", "<i4"), ("iterations", "<i4"), ("n_switch", "<i4"),
                         ("final_mismatch", "<f8"), ("vmin", "<f8"), ("vmax", "<f8"), ("max_loading", "<f8"),
                         ("n_voltage_violations", "<i4"), ("n_overloaded", "<i4"), ("island_count", "<i4"),
                         ("n_tiny_pivots", "<i4")])
assert RECORD_DTYPE.itemsize == RECORD_BYTES


def shard_range(total: int, rank: int, world: int):
    """Contiguous range [lo, hi) of scenarios for `rank` in input order like
    `Iterators.partition(eachindex(cases), ...)` (src/contingency.jl:304), balanced so that shard sizes differ by at
    most one - the same bounds as sblx_shard_bounds."""
    return total * rank // world, total * (rank + 1) // world


def pack_records(res: dict) -> np.ndarray:
    """host results dict (sblx.api.Engine.solve_outages) -> packed record array"""
    B = len(res["converged"])
    rec = np.zeros(B, RECORD_DTYPE)
    rec["converged"] = res["converged"]
    rec["status"] = res["status"]
    rec["iterations"] = res["iterations"]
    rec["n_switch"] = res["n_switch_events"]
    rec["final_mismatch"] = res["final_mismatch"]
    rec["vmin"] = res["vmin_pu"]
    rec["vmax"] = res["vmax_pu"]
    rec["max_loading"] = res["max_loading_pct"]
    for k in ("n_voltage_violations", "n_overloaded", "island_count", "n_tiny_pivots"):
        if k in res:
            rec[k] = res[k]
    return rec


def gather_records(local_u8, total: int, world: int, dist=None):
    """All-gather the byte tensor of local records (uint8, len = per_rank*48, padded to the common
    per-rank size) and return the records of all `total` scenarios in input order (torch uint8 tensor
    on the same device).  One collective."""
    import torch
    if dist is None:
        import torch.distributed as dist
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    per = max(max(sizes), 1)
    need = per * RECORD_BYTES
    if local_u8.numel() < need:
        pad = torch.zeros(need - local_u8.numel(), dtype=torch.uint8, device=local_u8.device)
        local_u8 = torch.cat([local_u8, pad])
    out = torch.empty(need * world, dtype=torch.uint8, device=local_u8.device)
    dist.all_gather_into_tensor(out, local_u8[:need].contiguous())
    parts = [out[r * need: r * need + sizes[r] * RECORD_BYTES] for r in range(world)]
    return torch.cat(parts)


def records_from_bytes(t) -> np.ndarray:
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=RECORD_DTYPE)
