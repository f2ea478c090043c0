"""Partition-independent 64-bit checksum of the OWNED results of a dynamics step.

The reference's regression policy across rank counts is bit-for-bit equality of the owned fields
(testing_and_setup/testing/tests/parallelism.py:75-85).  Gathering 10 M-cell fields to one rank just to compare
them is wasteful, so every rank hashes each owned value together with its GLOBAL id (and slot, and field name) and
the per-rank sums are added modulo 2**64: the result does not depend on how the entities are distributed or
ordered, and changes when any bit of any owned value, or the entity it belongs to, changes.
"""
from __future__ import annotations

import zlib

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    x = (x ^ (x >> np.uint64(30))) * _M1
    x = (x ^ (x >> np.uint64(27))) * _M2
    return x ^ (x >> np.uint64(31))


def field_checksum(name: str, values: np.ndarray, global_ids: np.ndarray, slots_valid: np.ndarray | None = None) -> int:
    """values: (n,) or (n, M) float64 of the owned entities; global_ids: (n,) 1-based ids; slots_valid: (n, M) bool."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    gid = np.asarray(global_ids).astype(np.uint64)
    salt = np.uint64(zlib.crc32(name.encode()))
    with np.errstate(over="ignore"):
        if v.ndim == 1:
            key = _mix(gid * _G + salt)
            h = _mix(key ^ v.view(np.uint64))
        else:
            m = v.shape[1]
            key = _mix((gid[:, None] * np.uint64(m) + np.arange(m, dtype=np.uint64)[None, :]) * _G + salt)
            h = _mix(key ^ v.view(np.uint64))
            if slots_valid is not None:
                h = np.where(slots_valid, h, np.uint64(0))
        return int(np.add.reduce(h.ravel(), dtype=np.uint64))


def owned_checksum(mesh, out: dict, vertex_fields=("uVelocity", "vVelocity"),
                   cell_fields=("stress11", "stress22", "stress12")) -> int:
    """Sum over the fields of field_checksum on the owned entities of ``mesh`` (a global mesh or a block of
    partition.build_block); add the values of all ranks modulo 2**64 to get the job's checksum."""
    nVs = int(mesh.get("nVerticesSolve", mesh["nVertices"]))
    nCs = int(mesh.get("nCellsSolve", mesh["nCells"]))
    vid = mesh["indexToVertexID"][:nVs] if "indexToVertexID" in mesh else np.arange(1, nVs + 1)
    cid = mesh["indexToCellID"][:nCs] if "indexToCellID" in mesh else np.arange(1, nCs + 1)
    total = 0
    for n in vertex_fields:
        total += field_checksum(n, out[n][:nVs], vid)
    if cell_fields:
        valid = np.arange(int(mesh["maxEdges"]))[None, :] < np.asarray(mesh["nEdgesOnCell"])[:nCs, None]
        for n in cell_fields:
            total += field_checksum(n, out[n][:nCs], cid, valid)
    return total & 0xFFFFFFFFFFFFFFFF


def combine(parts) -> int:
    return sum(int(p) for p in parts) & 0xFFFFFFFFFFFFFFFF
