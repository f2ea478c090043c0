"""The partition-independent checksum bench.py prints (mpas-seaice_b200/checksum.py): the sum over the blocks of a
decomposition equals the single-block value, whatever the partition, and any changed bit or any value attached to
the wrong entity changes it."""
import numpy as np
import pytest

import common
from mpas_seaice_b200 import checksum, partition


def _fields(mesh, seed=3):
    rng = np.random.default_rng(seed)
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    out = {k: rng.standard_normal(nV + 1) for k in ("uVelocity", "vVelocity")}
    out.update({k: rng.standard_normal((nC + 1, M)) for k in ("stress11", "stress22", "stress12")})
    return out


@pytest.mark.parametrize("kind,parts,method", [("ico3", 2, "auto"), ("ico3", 5, "rcb"), ("hex20", 3, "rcb"), ("quad40", 4, "rcb")])
def test_blocks_add_up_to_the_global_checksum(kind, parts, method):
    mesh, _ = common.mesh_case(kind)
    out = _fields(mesh)
    whole = checksum.owned_checksum(mesh, out)
    part = partition.partition_cells(mesh, parts, method)
    pieces = []
    for r in range(parts):
        blk = partition.build_block(mesh, part, r)
        local = partition.restrict_step(blk, out, mesh.nCells, mesh.nVertices)
        # halo and junk entries must not matter
        nVs, nCs = blk.nVerticesSolve, blk.nCellsSolve
        local["uVelocity"][nVs:] = 123.0
        local["stress11"][nCs:] = -7.0
        pieces.append(checksum.owned_checksum(blk, local))
    assert checksum.combine(pieces) == whole


def test_checksum_sees_single_bits_and_misplaced_values():
    mesh, _ = common.mesh_case("ico3")
    out = _fields(mesh)
    base = checksum.owned_checksum(mesh, out)
    flipped = {k: v.copy() for k, v in out.items()}
    flipped["stress12"][17, 2] = np.nextafter(flipped["stress12"][17, 2], np.inf)       # one bit
    assert checksum.owned_checksum(mesh, flipped) != base
    swapped = {k: v.copy() for k, v in out.items()}
    swapped["uVelocity"][[5, 6]] = swapped["uVelocity"][[6, 5]]                         # right values, wrong vertices
    assert checksum.owned_checksum(mesh, swapped) != base
    crossed = {k: v.copy() for k, v in out.items()}
    crossed["uVelocity"], crossed["vVelocity"] = crossed["vVelocity"], crossed["uVelocity"]      # right values, wrong field
    assert checksum.owned_checksum(mesh, crossed) != base
    # slots beyond nEdgesOnCell do not count
    padded = {k: v.copy() for k, v in out.items()}
    pent = np.nonzero(mesh.nEdgesOnCell[:mesh.nCells] == 5)[0]
    assert pent.size == 12
    padded["stress11"][pent, 5] = 99.0
    assert checksum.owned_checksum(mesh, padded) == base
