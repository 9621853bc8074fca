"""Scenario sharding across ranks (one process per GPU) and the single collective of the path:
an all-gather of fixed-size per-scenario result records (48 bytes: int32 converged, status,
iterations, n_switch; float64 final_mismatch, vmin, vmax, max_loading).

The reference fans contingencies out in contiguous chunks with no communication between workers
and returns results in input order (src/contingency.jl:299-318); ranks do the same with
`shard_range`, and `gather_records` restores the global input order.  Backend: NCCL on GPUs
(records stay in HBM: the tensor wraps the engine's device buffer), gloo in the CPU tests."""
from __future__ import annotations

import numpy as np

RECORD_BYTES = 48
RECORD_DTYPE = np.dtype([("converged", "<i4"), ("status", "<i4"), ("iterations", "<i4"), ("n_switch", "<i4"),
                         ("final_mismatch", "<f8"), ("vmin", "<f8"), ("vmax", "<f8"), ("max_loading", "<f8")])
assert RECORD_DTYPE.itemsize == RECORD_BYTES


def shard_range(total: int, rank: int, world: int):
    """Contiguous range [lo, hi) of scenarios for `rank`: ceil(total/world) per rank like
    `Iterators.partition(eachindex(cases), cld(n, cap))` (src/contingency.jl:304)."""
    per = -(-total // world)
    lo = min(total, rank * per)
    hi = min(total, lo + per)
    return lo, hi


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
    return rec


def gather_records(local_u8, total: int, world: int, dist=None):
    """All-gather the byte tensor of local records (uint8, len = per_rank*48, padded to the common
    per-rank size) and return the records of all `total` scenarios in input order (torch uint8 tensor
    on the same device).  One collective."""
    import torch
    if dist is None:
        import torch.distributed as dist
    per = -(-total // world)
    need = per * RECORD_BYTES
    if local_u8.numel() < need:
        pad = torch.zeros(need - local_u8.numel(), dtype=torch.uint8, device=local_u8.device)
        local_u8 = torch.cat([local_u8, pad])
    out = torch.empty(need * world, dtype=torch.uint8, device=local_u8.device)
    dist.all_gather_into_tensor(out, local_u8.contiguous())
    return out[: total * RECORD_BYTES]


def records_from_bytes(t) -> np.ndarray:
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=RECORD_DTYPE)
