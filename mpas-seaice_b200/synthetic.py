"""Synthetic ice states, analytic forcing and the host-side pre-subcycle (numpy, vectorised).

This plays the role of the Fortran host in front of the C-ABI: it produces, in the Registry's
layout, exactly the per-step fields ``velocity_solver_pre_subcycle`` hands to the subcycle
(reference: src/shared/mpas_seaice_velocity_solver.F:613-671).  Everything is closed-form and
seedless; no file is read.

The arithmetic order of every routine follows the cited reference lines so that tests can compare
these fields bit-for-bit with the C oracle's restatement of the same routines.
"""
from __future__ import annotations

import numpy as np

# reference: src/shared/mpas_seaice_constants.F:43-92, src/column/constants/cice/ice_constants_colpkg.F90:22-63
DENSITY_ICE = 917.0
DENSITY_SNOW = 330.0
HIBLER_P = 2.75e4
HIBLER_C = 20.0
AREA_MIN = 0.001      # velocity_solver.F:64
MASS_MIN = 0.01       # velocity_solver.F:65
AIR_STRESS_COEFF = 0.0012  # velocity_solver.F:1690


def time_steps(config_dt, n_dynamics=1, n_elastic=120):
    """dynamicsTimeStep, elasticTimeStep (velocity_solver.F:155-157), dampingTimescale
    (constitutive_relation.F:125)."""
    dt_dyn = config_dt / float(n_dynamics)
    dte = dt_dyn / float(n_elastic)
    return dt_dyn, dte, 0.36 * dt_dyn


def numerical_inertia_coefficient(dt_dyn, dv_edge_min):
    """constitutive_relation.F:154-162"""
    gamma = 0.25 * 1.0e11 * dt_dyn
    return (2.0 * 0.86 * 5.5e-3 * gamma) / (dv_edge_min * dv_edge_min)


# ---------------------------------------------------------------------------------------------
# states and forcing
# ---------------------------------------------------------------------------------------------

def square_state(mesh):
    """init_square_test_case_{state,atmos,ocean} (src/shared/mpas_seaice_testing.F:311-343,357-422,436-525),
    time = 0, Lx = Ly = 1.28e6."""
    Lx = Ly = 1.28e6
    x, y = mesh.xCell, mesh.yCell
    a, b = 5.0, 3.0
    theta = 4.0 * 24.0 * 3600.0
    time = 0.0
    s = np.sin((2.0 * np.pi * time) / theta) - b
    st = {}
    st["uAirVelocity"] = a + s * np.sin(2.0 * np.pi * (x / Lx)) * np.sin(np.pi * (y / Ly))
    st["vAirVelocity"] = a + s * np.sin(2.0 * np.pi * (y / Ly)) * np.sin(np.pi * (x / Lx))
    st["airDensity"] = np.full_like(x, 1.3)
    st["uOceanVelocity"] = 0.1 * ((2.0 * y - Ly) / Ly)
    st["vOceanVelocity"] = -0.1 * ((2.0 * x - Lx) / Lx)
    area = np.maximum(np.minimum(x / Lx, 1.0), 0.0)
    st["iceAreaCell"] = area
    st["iceVolumeCell"] = 2.0 * area
    st["snowVolumeCell"] = np.zeros_like(x)
    return st


def sphere_state(mesh, kind="A", perturb=True):
    """SURVEY.md section 8d configs 3-5.  kind 'A': full cover a=0.95, h=2 (every cell active);
    'B': caps lat>70N or lat<-60S with a=1, h=1 (reference cap recipe: mpas_seaice_initialize.F:530-537)."""
    lat, lon = mesh.latCell, mesh.lonCell
    st = {}
    ua = 10.0 * np.cos(lat)
    if perturb:
        ua = ua + 1.0e-3 * np.sin(7.0 * lon) * np.cos(5.0 * lat)
    st["uAirVelocity"] = ua
    st["vAirVelocity"] = 3.0 * np.sin(2.0 * lon) * np.cos(lat)
    st["airDensity"] = np.full_like(lat, 1.3)
    st["uOceanVelocity"] = 0.1 * np.cos(lat)
    st["vOceanVelocity"] = np.zeros_like(lat)
    if kind == "A":
        area = np.full_like(lat, 0.95)
        thick = 2.0
    elif kind == "B":
        cap = (lat > np.deg2rad(70.0)) | (lat < np.deg2rad(-60.0))
        area = np.where(cap, 1.0, 0.0)
        thick = 1.0
    else:
        raise ValueError(kind)
    st["iceAreaCell"] = area
    st["iceVolumeCell"] = thick * area
    st["snowVolumeCell"] = np.zeros_like(lat)
    return st


def cell_inputs(state):
    """The CELL fields a host hands to evp_pre_subcycle (include/evp_b200.h, evp_pre_fields): what stays on
    the host of velocity_solver_pre_subcycle -- aggregate_mass_and_area for a single category
    (velocity_solver.F:738-746) and the Hibler ice strength before its solveStress mask (:1419-1436; exp()
    is host-side on purpose) -- plus the atmosphere / ocean coupling fields as they are."""
    area = np.ascontiguousarray(state["iceAreaCell"], dtype=np.float64)
    vol = np.ascontiguousarray(state["iceVolumeCell"], dtype=np.float64)
    snow = np.ascontiguousarray(state["snowVolumeCell"], dtype=np.float64)
    c = dict(iceAreaCellInitial=area, iceAreaCell=area,
             totalMassCell=vol * DENSITY_ICE + snow * DENSITY_SNOW,
             icePressure=HIBLER_P * vol * np.exp(-HIBLER_C * (1.0 - area)))
    for k in ("uOceanVelocity", "vOceanVelocity", "uAirVelocity", "vAirVelocity", "airDensity"):
        c[k] = np.ascontiguousarray(state[k], dtype=np.float64)
    return c


# ---------------------------------------------------------------------------------------------
# pre-subcycle
# ---------------------------------------------------------------------------------------------

def interpolate_cell_to_vertex(mesh, var_cell, n_vertices_solve=None):
    """seaice_interpolate_cell_to_vertex, cell-area weights, NO validity test on the cell index
    (src/shared/mpas_seaice_mesh.F:2835-2851): boundary vertices get junk by design."""
    nV = mesh.nVertices
    nVs = nV if n_vertices_solve is None else n_vertices_solve
    cov = mesh.cellsOnVertex[:nVs] - 1
    acc = np.zeros(nVs)
    tot = np.zeros(nVs)
    with np.errstate(all="ignore"):
        for k in range(mesh.vertexDegree):
            c = cov[:, k]
            acc = acc + mesh.areaCell[c] * var_cell[c]
            tot = tot + mesh.areaCell[c]
        out = np.zeros(nV + 1)
        out[:nVs] = acc / tot
    return out


def interior_vertex(mesh):
    """interior_vertices (mesh.F:423-488)"""
    out = np.zeros(mesh.nVertices + 1, dtype=np.int32)
    cov = mesh.cellsOnVertex[:mesh.nVertices]
    out[:mesh.nVertices] = np.all((cov >= 1) & (cov <= mesh.nCells), axis=1)
    return out


def pre_subcycle(mesh, state, config_dt, *, n_elastic=120, prev=None, use_air_stress=True,
                 use_ocean_stress=True, use_surface_tilt=True, constitutive_relation_type="evp",
                 masks=None, land_ice_mask=None):
    """velocity_solver_pre_subcycle (velocity_solver.F:613-671) for a single-category state without
    the column package.  ``prev`` carries the state that survives between dynamics steps
    (uVelocity, vVelocity, stress11/22/12, solveVelocityPrevious); None = cold start.
    ``masks`` = (solveStress, solveVelocity) overrides calculation_masks, which is what
    config_calc_velocity_masks = false does for the operator tests (velocity_solver.F:897-901).
    ``land_ice_mask`` (nCells+1): cells under an ice shelf, excluded from both masks (:1023, :1131 through
    init_ice_shelve_vertex_mask :481-544); None = none.
    Returns (step_fields, options)."""
    nC, nV, M = mesh.nCells, mesh.nVertices, mesh.maxEdges
    dt_dyn, dte, damping = time_steps(config_dt, 1, n_elastic)
    f = {}

    # aggregate_mass_and_area (:685-752)
    ice_area = np.array(state["iceAreaCell"], dtype=np.float64)
    ice_vol = np.array(state["iceVolumeCell"], dtype=np.float64)
    snow_vol = np.array(state["snowVolumeCell"], dtype=np.float64)
    total_mass_cell = ice_vol * DENSITY_ICE + snow_vol * DENSITY_SNOW

    # calculation_masks (:766-947)
    f["iceAreaVertex"] = interpolate_cell_to_vertex(mesh, ice_area)
    f["totalMassVertex"] = interpolate_cell_to_vertex(mesh, total_mass_cell)
    land_ice = np.zeros(nC + 1, dtype=np.int32) if land_ice_mask is None else np.asarray(land_ice_mask, dtype=np.int32)
    if masks is None:
        # stress_calculation_mask (:961-1059)
        enough = (ice_area > AREA_MIN) & (total_mass_cell > MASS_MIN) & (land_ice == 0)
        enough[nC] = False
        nb_enough = np.zeros(nC + 1, dtype=bool)
        for k in range(M):
            valid = mesh.nEdgesOnCell > k
            nb_enough |= valid & enough[mesh.cellsOnCell[:, k] - 1]
        solve_stress = (enough | nb_enough).astype(np.int32)
        solve_stress[nC] = 0
        # velocity_calculation_mask (:1073-1150)
        interior = interior_vertex(mesh)
        with np.errstate(invalid="ignore"):
            land_ice_vertex = np.zeros(nV + 1, dtype=bool)
            land_ice_vertex[:nV] = np.any(land_ice[mesh.cellsOnVertex[:nV] - 1] == 1, axis=1)
            solve_velocity = ((interior == 1) & ~land_ice_vertex & (f["iceAreaVertex"] > AREA_MIN) &
                              (f["totalMassVertex"] > MASS_MIN)).astype(np.int32)
        solve_velocity[nV] = 0
    else:
        solve_stress = np.array(masks[0], dtype=np.int32)
        solve_velocity = np.array(masks[1], dtype=np.int32)
    f["solveStress"] = solve_stress
    f["solveVelocity"] = solve_velocity
    sv = solve_velocity == 1

    # new_ice_velocities (:1164-1327)
    f["uOceanVelocityVertex"] = interpolate_cell_to_vertex(mesh, state["uOceanVelocity"])
    f["vOceanVelocityVertex"] = interpolate_cell_to_vertex(mesh, state["vOceanVelocity"])
    if prev is None:
        u = np.zeros(nV + 1)
        v = np.zeros(nV + 1)
        sv_prev = solve_velocity.copy()   # cold start from rest (config_initial_velocity_type = none)
        s11 = np.zeros((nC + 1, M))
        s22 = np.zeros((nC + 1, M))
        s12 = np.zeros((nC + 1, M))
    else:
        u, v = prev["uVelocity"].copy(), prev["vVelocity"].copy()
        sv_prev = prev["solveVelocityPrevious"]
        s11, s22, s12 = prev["stress11"].copy(), prev["stress22"].copy(), prev["stress12"].copy()
    new_ice = sv & (sv_prev == 0)
    u[new_ice] = f["uOceanVelocityVertex"][new_ice]
    v[new_ice] = f["vOceanVelocityVertex"][new_ice]
    u[~sv] = 0.0
    v[~sv] = 0.0
    f["solveVelocityPrevious"] = solve_velocity.copy()
    f["uVelocityInitial"] = u.copy()
    f["vVelocityInitial"] = v.copy()

    # ice_strength, Hibler branch (:1419-1436)
    f["icePressure"] = np.where(solve_stress == 1, HIBLER_P * ice_vol * np.exp(-HIBLER_C * (1.0 - ice_area)), 0.0)

    # air_stress -> constant_air_stress (:1665-1728) + interpolation (:1635-1650)
    if use_air_stress:
        ua, va, rho = state["uAirVelocity"], state["vAirVelocity"], state["airDensity"]
        wind = np.sqrt(ua * ua + va * va)
        air_u = rho * wind * AIR_STRESS_COEFF * ua * ice_area
        air_v = rho * wind * AIR_STRESS_COEFF * va * ice_area
    else:
        air_u = np.zeros(nC + 1)
        air_v = np.zeros(nC + 1)
    f["airStressVertexU"] = interpolate_cell_to_vertex(mesh, air_u)
    f["airStressVertexV"] = interpolate_cell_to_vertex(mesh, air_v)

    # coriolis_force_coefficient (:1742-1788)
    with np.errstate(invalid="ignore"):
        f["totalMassVertexfVertex"] = f["totalMassVertex"] * mesh.fVertex

        # ocean_stress (:1802-1883): turning angle 0 => cos = 1, sin = 0, terms kept
        sgn = np.copysign(1.0, mesh.fVertex)
        if use_ocean_stress:
            ou = f["uOceanVelocityVertex"] * 1.0 - f["vOceanVelocityVertex"] * 0.0 * sgn
            ov = f["uOceanVelocityVertex"] * 0.0 * sgn + f["vOceanVelocityVertex"] * 1.0
            f["oceanStressU"] = np.where(sv, ou, 0.0)
            f["oceanStressV"] = np.where(sv, ov, 0.0)
        else:
            f["oceanStressU"] = np.zeros(nV + 1)
            f["oceanStressV"] = np.zeros(nV + 1)

        # surface_tilt_geostrophic (:1941-2010) | no_surface_tilt
        if use_surface_tilt:
            f["surfaceTiltForceU"] = np.where(sv, -mesh.fVertex * f["totalMassVertex"] * f["vOceanVelocityVertex"], 0.0)
            f["surfaceTiltForceV"] = np.where(sv, mesh.fVertex * f["totalMassVertex"] * f["uOceanVelocityVertex"], 0.0)
        else:
            f["surfaceTiltForceU"] = np.zeros(nV + 1)
            f["surfaceTiltForceV"] = np.zeros(nV + 1)

    # init_subcycle_variables (:2227-2386)
    f["stressDivergenceU"] = np.zeros(nV + 1)
    f["stressDivergenceV"] = np.zeros(nV + 1)
    f["oceanStressCoeff"] = np.zeros(nV + 1)
    f["uVelocity"] = u
    f["vVelocity"] = v
    f["strain11"] = np.zeros((nC + 1, M))
    f["strain22"] = np.zeros((nC + 1, M))
    f["strain12"] = np.zeros((nC + 1, M))
    off = solve_stress != 1
    s11[off] = 0.0
    s22[off] = 0.0
    s12[off] = 0.0
    f["stress11"], f["stress22"], f["stress12"] = s11, s22, s12
    f["replacementPressure"] = np.zeros((nC + 1, M))

    # junk produced at non-interior vertices by the unguarded interpolation (NaN/huge) is masked by
    # solveVelocity in every consumer; scrub it so host<->device comparisons of inputs are finite.
    for k in ("iceAreaVertex", "totalMassVertex", "uOceanVelocityVertex", "vOceanVelocityVertex",
              "airStressVertexU", "airStressVertexV", "totalMassVertexfVertex"):
        bad = ~np.isfinite(f[k])
        f[k][bad] = 0.0

    opts = dict(constitutive_relation_type=constitutive_relation_type, ocean_stress_type="quadratic",
                use_ocean_stress=use_ocean_stress, average_variational_strain=False,
                elasticTimeStep=dte, dynamicsTimeStep=dt_dyn, dampingTimescale=damping,
                numericalInertiaCoefficient=numerical_inertia_coefficient(dt_dyn, float(mesh.dvEdge[:-1].min())),
                n_elastic=n_elastic)
    return f, opts


def operator_test_fields(mesh, A=2.56, B=2.56, Cc=2.56, Dd=2.56):
    """The analytic velocity / strain / stress-divergence of the reference's operator test
    (testing_and_setup/testcases/square/operators_strain_stress_divergence/create_ics.py:12-48),
    evaluated at vertices with x, y measured from the domain minimum (:96-97)."""
    Lx = Ly = 1.0
    x = mesh.xVertex - mesh.xVertex[:-1].min()
    y = mesh.yVertex - mesh.yVertex[:-1].min()
    pi = np.pi
    sx, cx = np.sin((2 * pi * x * A) / Lx), np.cos((2 * pi * x * A) / Lx)
    sy, cy = np.sin((2 * pi * y * B) / Ly), np.cos((2 * pi * y * B) / Ly)
    sx2, cx2 = np.sin((2 * pi * x * Cc) / Lx), np.cos((2 * pi * x * Cc) / Lx)
    sy2, cy2 = np.sin((2 * pi * y * Dd) / Ly), np.cos((2 * pi * y * Dd) / Ly)
    kA, kB, kC, kD = (2 * pi * A) / Lx, (2 * pi * B) / Ly, (2 * pi * Cc) / Lx, (2 * pi * Dd) / Ly
    u = sx * sy
    v = sx2 * sy2
    dudx, dudy = kA * cx * sy, kB * sx * cy
    dvdx, dvdy = kC * cx2 * sy2, kD * sx2 * cy2
    d2udx2, d2udy2, d2udxdy = -kA * kA * sx * sy, -kB * kB * sx * sy, kA * kB * cx * cy
    d2vdx2, d2vdy2, d2vdxdy = -kC * kC * sx2 * sy2, -kD * kD * sx2 * sy2, kC * kD * cx2 * cy2
    e11, e22, e12 = dudx, dvdy, 0.5 * (dudy + dvdx)
    divu = d2udx2 + 0.5 * (d2udy2 + d2vdxdy)
    divv = 0.5 * (d2udxdy + d2vdx2) + d2vdy2
    return dict(u=u, v=v, e11=e11, e22=e22, e12=e12, divu=divu, divv=divv)
