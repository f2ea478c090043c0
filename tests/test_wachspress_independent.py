"""An independent check of the Wachspress precompute (oracle/evp_precompute_oracle.c, device csrc/evp_precompute.cu):
the basis functions are rebuilt here from a DIFFERENT closed form than the reference's edge-line products with the
kappa recursion (src/shared/mpas_seaice_velocity_solver_wachspress.F:573-668) -- the triangle-area form
    w_i(x) = C_i * prod_{k not in {i-1, i}} A_k(x),   A_k(x) = area(x, v_k, v_{k+1}),  C_i = area(v_{i-1}, v_i, v_{i+1}),
    phi_i = w_i / sum_j w_j
(Meyer et al. 2002; Floater 2015) -- differentiated analytically, and the integrals over the cell are taken with a
tensor Gauss rule of 24 x 24 points per sub-triangle instead of the reference's 16-point Dunavant rule.  What must agree:
  * basisGradientU/V (values at the cell's vertices): to round-off -- same functions, no quadrature involved;
  * basisIntegralsU/V/Metric: to the truncation error of the Dunavant order-8 rule on these rational functions.
This pins the mathematics of the restated precompute (which function sits in which (i, j) slot included); it cannot
pin the reference's choice of quadrature constants, which are compared digit for digit in the source instead."""
import numpy as np
import pytest

import common


def _area(p, a, b):
    return 0.5 * ((a[0] - p[0]) * (b[1] - p[1]) - (a[1] - p[1]) * (b[0] - p[0]))


def _grad_area(a, b):
    """gradient with respect to p of area(p, a, b)"""
    return 0.5 * np.array([a[1] - b[1], b[0] - a[0]])


def _phi_and_grad(verts, p):
    n = len(verts)
    A = np.array([_area(p, verts[k], verts[(k + 1) % n]) for k in range(n)])
    gA = np.array([_grad_area(verts[k], verts[(k + 1) % n]) for k in range(n)])
    w = np.zeros(n)
    gw = np.zeros((n, 2))
    for i in range(n):
        C = _area(verts[(i - 1) % n], verts[i], verts[(i + 1) % n])
        ks = [k for k in range(n) if k not in ((i - 1) % n, i)]
        w[i] = C * np.prod(A[ks])
        for k in ks:
            gw[i] += C * np.prod(A[[m for m in ks if m != k]]) * gA[k]
    S = w.sum()
    phi = w / S
    gphi = gw / S - np.outer(w, gw.sum(axis=0)) / S ** 2
    return phi, gphi


def _gauss_triangle(n):
    """points (u, v) and weights on the unit triangle u, v >= 0, u + v <= 1 (Duffy transform of an n x n Gauss rule)"""
    x, wx = np.polynomial.legendre.leggauss(n)
    x, wx = 0.5 * (x + 1.0), 0.5 * wx
    U, V, W = [], [], []
    for a, wa in zip(x, wx):
        for b, wb in zip(x, wx):
            U.append(a)
            V.append(b * (1.0 - a))
            W.append(wa * wb * (1.0 - a))
    return np.array(U), np.array(V), np.array(W)


def _reference_integrals(verts):
    n = len(verts)
    U, V, W = _gauss_triangle(24)
    IU, IV, IM = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for s in range(n):
        a, b = verts[s], verts[(s + 1) % n]
        jac = a[0] * b[1] - b[0] * a[1]                  # sub-triangle (0, v_s, v_{s+1}): x = a u + b v
        for u, v, wq in zip(U, V, W):
            p = a * u + b * v
            phi, g = _phi_and_grad(verts, p)
            IU += jac * wq * np.outer(phi, g[:, 0])      # [i, j] = phi_i dphi_j/dx
            IV += jac * wq * np.outer(phi, g[:, 1])
            IM += jac * wq * np.outer(phi, phi)
    return IU, IV, IM


CASES = [("hex20", "interior hexagon"), ("quad40", "interior square"), ("ico3", "pentagon"), ("ico3", "irregular hexagon")]


def _pick_cell(mesh, what):
    nC = mesh.nCells
    n = mesh.nEdgesOnCell[:nC]
    if what == "pentagon":
        return int(np.nonzero(n == 5)[0][0])
    if what == "irregular hexagon":
        pent = np.nonzero(n == 5)[0][0]
        return int(mesh.cellsOnCell[pent, 0] - 1)          # a hexagon next to a pentagon: the most distorted ones
    interior = np.all(mesh.cellsOnCell[:nC, :n.max()] <= nC, axis=1) & (n == n.max())
    return int(np.nonzero(interior)[0][len(np.nonzero(interior)[0]) // 2])


@pytest.mark.parametrize("kind,what", CASES)
def test_gradients_and_integrals_against_the_area_form(kind, what):
    mesh, var = common.mesh_case(kind, metric=False)
    c = _pick_cell(mesh, what)
    n = int(mesh.nEdgesOnCell[c])
    verts = np.stack([var["xLocal"][c, :n], var["yLocal"][c, :n]], axis=1)
    scale = np.abs(verts).max()
    # ---- gradients at the vertices: numpy [c, j, i] = dphi_i at vertex j ----
    GU, GV = var["basisGradientU"][c, :n, :n], var["basisGradientV"][c, :n, :n]
    for j in range(n):
        _, g = _phi_and_grad(verts, verts[j])
        assert np.abs(GU[j] - g[:, 0]).max() <= 1e-11 / scale, (j, GU[j], g[:, 0])
        assert np.abs(GV[j] - g[:, 1]).max() <= 1e-11 / scale
    # ---- integrals: numpy [c, j, i] = integral of phi_i dphi_j/dx ----
    IU, IV, IM = _reference_integrals(verts)
    SU, SV, SM = (var[k][c, :n, :n] for k in ("basisIntegralsU", "basisIntegralsV", "basisIntegralsMetric"))
    # truncation of the 16-point Dunavant rule on these rational functions: < 2e-5 on regular cells, 8e-5 measured on the
    # most distorted hexagon of the coarsest sphere (next to a pentagon of the 642-cell mesh)
    tol = 2e-4 if what == "irregular hexagon" else 2e-5
    assert np.abs(SU - IU.T).max() <= tol * np.abs(IU).max(), np.abs(SU - IU.T).max() / np.abs(IU).max()
    assert np.abs(SV - IV.T).max() <= tol * np.abs(IV).max()
    assert np.abs(SM - IM.T).max() <= tol * np.abs(IM).max()
    # the slot convention is not symmetric: the transposed reading must fail clearly for U and V
    if what != "interior square":
        assert np.abs(SU - IU).max() > 100 * tol * np.abs(IU).max()
    # exact identities of the high-order reference itself (a check of this file, not of the oracle)
    area = 0.5 * sum(verts[s][0] * verts[(s + 1) % n][1] - verts[(s + 1) % n][0] * verts[s][1] for s in range(n))
    assert abs(IM.sum() - area) <= 1e-12 * area
    assert np.abs(IU.sum(axis=1)).max() <= 1e-12 * np.abs(IU).max()       # sum_j dphi_j/dx = 0
