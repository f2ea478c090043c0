"""B200-native EVP momentum subcycle for MPAS-Seaice (drop-in for the body of
``subcycle_velocity_solver``, reference: src/shared/mpas_seaice_velocity_solver.F:2404-2464).

Layout:
  csrc/        CUDA kernels (sm_100a) + the C-ABI shared library ``libevp_b200.so``
  host.py      ctypes mirror of the C-ABI, named after the reference's seaice_mesh_pool lifecycle
  meshgen.py   offline mesh generators (planar hex / quad, icosahedral sphere) in MPAS conventions
  synthetic.py synthetic ice states + analytic forcing + the host-side pre-subcycle fields
  partition.py MPAS-style graph decomposition, halo layers and exchange lists
"""
__version__ = "0.1.0"
