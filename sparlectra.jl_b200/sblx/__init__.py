"""sblx - host side of the B200-native batched AC power-flow engine (Sparlectra hot path).

`sblx.grid`   neutral grid tables + case generators
`sblx.model`  grid tables -> solver arrays (Ybus, bus types, S, Vset, Q limits, branch stamps)
`sblx.lib`    ctypes binding of libsblx.so (include/sblx.h)
`sblx.api`    reference-shaped interface: PFModel/PFSolution/solvePf, runContingenciesBatched
"""
from . import grid, model  # noqa: F401
