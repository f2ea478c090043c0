"""B200-native EVP momentum subcycle for MPAS-Seaice (drop-in for the body of
``subcycle_velocity_solver``, reference: src/shared/mpas_seaice_velocity_solver.F:2404-2464).

Layout:
  csrc/               CUDA kernels (sm_100a) + the C-ABI shared library ``libevp_b200.so``
                      (evp_kernels.cu: subcycle; evp_prepost.cu: pre-/post-subcycle; evp_weak.cu: weak operators;
                      evp_precompute.cu: Wachspress / PWL basis; evp_halo.cu: NCCL exchange; evp_abi.cu: lifecycle)
  host.py             ctypes mirror of the C-ABI, named after the reference's seaice_mesh_pool lifecycle
  meshgen.py          offline mesh generators (planar hex / quad, icosahedral sphere) in MPAS conventions
  variational_init.py host-side static fields (local coordinates, metric terms, cellVerticesAtVertex, ...)
  weakmesh.py         edge connectivity + normal vectors for the weak operators (synthetic hosts)
  synthetic.py        synthetic ice states, analytic forcing, the host-side pre-subcycle, cell inputs
  partition.py        MPAS-style graph decomposition, halo layers and exchange lists
  multigpu.py         one block per rank: what a decomposed host hands to the C-ABI
  workloads.py        the BASELINE.json configurations as ready-to-run workloads
"""
__version__ = "0.1.0"
