"""MATPOWER `.m` case reader so the engine can be fed from real case files.

Follows the reference's text rules (src/MatpowerIO.jl): `%` comments are stripped to the end
of the line (:344-365); `mpc.baseMVA = <number>;` (:367-371); a matrix block is everything
between `<key> = [` and the next `];` (:391-421); rows are separated by `;`, CR or LF and
blank rows are skipped (:427, :495-514); `mpc.bus`, `mpc.gen`, `mpc.branch` are padded /
truncated with zeros to the MATPOWER v2 widths 13 / 21 / 13 (:278-280); bus rows are sorted
by BUS_I when they are not already (`legacy_sort_bus`, :174-182); raw TAP values are kept
(TAP == 0 = line, :184-189).  Only what the batched solver needs is read: no gencost,
dcline, name vectors or script post-processing."""
from __future__ import annotations

import os
import re
import numpy as np

from . import grid as G

_WIDTH = {"mpc.bus": 13, "mpc.gen": 21, "mpc.branch": 13}


def strip_matlab_comments(txt: str) -> str:
    return "\n".join(line.split("%", 1)[0] for line in txt.split("\n"))


def parse_base_mva(txt: str) -> float:
    m = re.search(r"mpc\.baseMVA\s*=\s*([0-9]+(?:\.[0-9]+)?)\s*;", txt)
    if m is None:
        raise ValueError("Could not find `mpc.baseMVA = ...;` in MATPOWER file.")
    return float(m.group(1))


def matrix_body(txt: str, key: str) -> str:
    k = txt.find(key)
    if k < 0:
        raise ValueError(f"Could not find matrix block `{key} = [ ... ];`")
    rest = txt[k + len(key):]
    m = re.match(r"\s*=\s*\[", rest)
    if m is None:
        raise ValueError(f"Could not find `= [` after matrix block key `{key}`.")
    start = k + len(key) + m.end()
    end = txt.find("];", start)
    if end < 0:
        raise ValueError(f"Could not find closing `];` for matrix block `{key}`.")
    return txt[start:end]


def parse_matrix(body: str, key: str, ncols: int) -> np.ndarray:
    rows = []
    for rr in re.split(r"[;\r\n]", body):
        toks = rr.split()
        if not toks:
            continue
        vals = [float(t) for t in toks[:ncols]]
        rows.append(vals + [0.0] * (ncols - len(vals)))
    if not rows:
        raise ValueError(f"Matrix `{key}` appears empty or could not be parsed.")
    return np.asarray(rows, dtype=np.float64)


def read_case_m(path_or_text: str, legacy_compat: bool = True) -> dict:
    """-> {"name", "baseMVA", "bus" [nb x 13], "gen" [ng x 21], "branch" [nl x 13]}"""
    if "\n" not in path_or_text and os.path.exists(path_or_text):
        name = os.path.splitext(os.path.basename(path_or_text))[0]
        with open(path_or_text, "r") as f:
            txt = f.read()
    else:
        txt = path_or_text
        m = re.search(r"function\s+mpc\s*=\s*(\w+)", txt)
        name = m.group(1) if m else "case"
    txt = strip_matlab_comments(txt)
    out = {"name": name, "baseMVA": parse_base_mva(txt)}
    for key, w in _WIDTH.items():
        out[key.split(".")[1]] = parse_matrix(matrix_body(txt, key), key, w)
    bus = out["bus"]
    if legacy_compat and not np.all(np.diff(bus[:, 0]) >= 0):
        out["bus"] = bus[np.argsort(bus[:, 0], kind="stable")]
    return out


def grid_from_case(path_or_text: str) -> G.Grid:
    c = read_case_m(path_or_text)
    return G.from_matpower_tables(c["name"], c["baseMVA"], c["bus"], c["gen"], c["branch"])


def read_case_m_native(path_or_text: str) -> dict:
    """The same tables through the library's C reader (sblx_mpc_read / sblx_mpc_parse, csrc/matpower.cpp)."""
    import ctypes as C
    from . import lib as L
    lib = L.load()
    h = C.c_void_p()
    if "\n" not in path_or_text and os.path.exists(path_or_text):
        name = os.path.splitext(os.path.basename(path_or_text))[0]
        rc = lib.sblx_mpc_read(path_or_text.encode(), C.byref(h))
    else:
        m = re.search(r"function\s+mpc\s*=\s*(\w+)", path_or_text)
        name = m.group(1) if m else "case"
        rc = lib.sblx_mpc_parse(path_or_text.encode(), C.byref(h))
    if rc != 0:
        raise ValueError(lib.sblx_mpc_last_error().decode())
    try:
        base = C.c_double()
        nb, ng, nl = C.c_int64(), C.c_int64(), C.c_int64()
        lib.sblx_mpc_dims(h, C.byref(base), C.byref(nb), C.byref(ng), C.byref(nl))
        bus, gen, br = np.zeros((nb.value, 13)), np.zeros((ng.value, 21)), np.zeros((nl.value, 13))
        lib.sblx_mpc_tables(h, L.ptr(bus, L.c_f64p), L.ptr(gen, L.c_f64p), L.ptr(br, L.c_f64p))
    finally:
        lib.sblx_mpc_free(h)
    return {"name": name, "baseMVA": base.value, "bus": bus, "gen": gen, "branch": br}
