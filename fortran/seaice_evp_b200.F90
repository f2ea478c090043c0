!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!
!  seaice_evp_b200
!
!> \brief ISO_C_BINDING shim between MPAS-Seaice and libevp_b200.so (include/evp_b200.h)
!>
!> Drop-in for the body of subcycle_velocity_solver
!> (src/shared/mpas_seaice_velocity_solver.F:2404-2464) and for the device-residency lifecycle of
!> module seaice_mesh_pool (src/shared/mpas_seaice_mesh_pool.F:76-281):
!>
!>    seaice_mesh_pool_create  ->  seaice_evp_b200_create   (after seaice_init_velocity_solver,
!>                                                           src/shared/mpas_seaice_initialize.F:121)
!>    seaice_mesh_pool_update  ->  seaice_evp_b200_update   (end of velocity_solver_pre_subcycle,
!>                                                           velocity_solver.F:668)
!>    subcycle_velocity_solver ->  seaice_evp_b200_subcycle (velocity_solver.F:585)
!>    seaice_mesh_pool_destroy ->  seaice_evp_b200_destroy  (src/model_forward/mpas_seaice_core.F:431)
!>
!> seaice_run_velocity_solver(domain, clock) keeps its signature, its namelist options and its
!> Registry fields.  Combinations the CUDA path does not cover (weak strain / weak divergence,
!> config_average_variational_strain) are reported by seaice_evp_b200_supported() = .false. and the
!> caller keeps using the original Fortran subcycle -- the ORIGINAL code, not a CPU copy of this one.
!>
!> The bind(C) types below mirror include/evp_b200.h field for field (tests/test_fortran_shim.py
!> checks names, order and types against the header).  No Fortran compiler exists in the build
!> container; the module is executed there by an interpreter that calls the library through these
!> interface blocks (tests/test_fortran_shim_executed.py); see INTEGRATION.md.
!
!-----------------------------------------------------------------------

module seaice_evp_b200

  use, intrinsic :: iso_c_binding

#ifndef EVP_B200_STANDALONE
  use mpass_derived_types
  use mpass_pool_routines
  use mpass_log, only: mpas_log_write
#endif

  implicit none

  private
  save

#ifndef EVP_B200_STANDALONE
  public :: &
       seaice_evp_b200_supported, &
       seaice_evp_b200_create, &
       seaice_evp_b200_update, &
       seaice_evp_b200_subcycle, &
       seaice_evp_b200_set_mesh_ext, &
       seaice_evp_b200_set_weak_mesh, &
       seaice_evp_b200_step, &
       seaice_evp_b200_destroy
#endif

  ! ---- enums of include/evp_b200.h ----
  integer(c_int), parameter, public :: &
       EVP_OK = 0, &
       EVP_CR_EVP = 1, EVP_CR_EVP_REVISED = 2, EVP_CR_LINEAR = 3, EVP_CR_NONE = 4, &
       EVP_OCEAN_QUADRATIC = 1, EVP_OCEAN_LINEAR = 2, &
       EVP_FLAG_PIN_HOST = 1, EVP_FLAG_OVERLAP_HALO = 2, &
       EVP_SCHEME_VARIATIONAL = 1, EVP_SCHEME_WEAK = 2, &
       EVP_START_RESIDENT = 0, EVP_START_FROM_REST = 1, EVP_START_FIRST_STEP = 2, &
       EVP_HALO_NONE = 0, EVP_HALO_NCCL = 1, EVP_HALO_P2P = 2

  ! ---- struct evp_mesh_desc ----
  type, bind(C), public :: evp_mesh_desc
     integer(c_int) :: nCells
     integer(c_int) :: nCellsSolve
     integer(c_int) :: nVertices
     integer(c_int) :: nVerticesSolve
     integer(c_int) :: maxEdges
     integer(c_int) :: vertexDegree
     type(c_ptr) :: nEdgesOnCell
     type(c_ptr) :: verticesOnCell
     type(c_ptr) :: cellsOnVertex
     type(c_ptr) :: cellVerticesAtVertex
     type(c_ptr) :: basisGradientU
     type(c_ptr) :: basisGradientV
     type(c_ptr) :: basisIntegralsU
     type(c_ptr) :: basisIntegralsV
     type(c_ptr) :: basisIntegralsMetric
     type(c_ptr) :: tanLatVertexRotatedOverRadius
     type(c_ptr) :: variationalDenominator
     type(c_ptr) :: vertexBoundaryType
     type(c_ptr) :: vertexBoundarySourceLocal
  end type evp_mesh_desc

  ! ---- struct evp_options ----
  type, bind(C), public :: evp_options
     integer(c_int) :: constitutive_relation_type
     integer(c_int) :: ocean_stress_type
     integer(c_int) :: use_ocean_stress
     integer(c_int) :: use_special_boundaries_velocity
     integer(c_int) :: device
     integer(c_int) :: flags
     integer(c_int) :: average_variational_strain
     integer(c_int) :: strain_scheme
     integer(c_int) :: stress_divergence_scheme
     real(c_double) :: elasticTimeStep
     real(c_double) :: dynamicsTimeStep
     real(c_double) :: dampingTimescale
     real(c_double) :: numericalInertiaCoefficient
  end type evp_options

  ! ---- struct evp_step_fields ----
  type, bind(C), public :: evp_step_fields
     type(c_ptr) :: solveStress
     type(c_ptr) :: solveVelocity
     type(c_ptr) :: icePressure
     type(c_ptr) :: uVelocity
     type(c_ptr) :: vVelocity
     type(c_ptr) :: stress11
     type(c_ptr) :: stress22
     type(c_ptr) :: stress12
     type(c_ptr) :: totalMassVertex
     type(c_ptr) :: totalMassVertexfVertex
     type(c_ptr) :: iceAreaVertex
     type(c_ptr) :: airStressVertexU
     type(c_ptr) :: airStressVertexV
     type(c_ptr) :: surfaceTiltForceU
     type(c_ptr) :: surfaceTiltForceV
     type(c_ptr) :: oceanStressU
     type(c_ptr) :: oceanStressV
     type(c_ptr) :: uOceanVelocityVertex
     type(c_ptr) :: vOceanVelocityVertex
     type(c_ptr) :: uVelocityInitial
     type(c_ptr) :: vVelocityInitial
  end type evp_step_fields

  ! ---- struct evp_out_fields ----
  type, bind(C), public :: evp_out_fields
     type(c_ptr) :: uVelocity
     type(c_ptr) :: vVelocity
     type(c_ptr) :: stress11
     type(c_ptr) :: stress22
     type(c_ptr) :: stress12
     type(c_ptr) :: strain11
     type(c_ptr) :: strain22
     type(c_ptr) :: strain12
     type(c_ptr) :: replacementPressure
     type(c_ptr) :: stressDivergenceU
     type(c_ptr) :: stressDivergenceV
     type(c_ptr) :: oceanStressCoeff
  end type evp_out_fields

  ! ---- struct evp_mesh_ext ----
  type, bind(C), public :: evp_mesh_ext
     type(c_ptr) :: cellsOnCell
     type(c_ptr) :: interiorVertex
     type(c_ptr) :: landIceMaskVertex
     type(c_ptr) :: areaCell
     type(c_ptr) :: areaTriangle
     type(c_ptr) :: fVertex
  end type evp_mesh_ext

  ! ---- struct evp_pre_fields ----
  type, bind(C), public :: evp_pre_fields
     type(c_ptr) :: iceAreaCellInitial
     type(c_ptr) :: iceAreaCell
     type(c_ptr) :: totalMassCell
     type(c_ptr) :: icePressure
     type(c_ptr) :: uOceanVelocity
     type(c_ptr) :: vOceanVelocity
     type(c_ptr) :: airStressCellU
     type(c_ptr) :: airStressCellV
     type(c_ptr) :: uAirVelocity
     type(c_ptr) :: vAirVelocity
     type(c_ptr) :: airDensity
     type(c_ptr) :: seaSurfaceTiltU
     type(c_ptr) :: seaSurfaceTiltV
     type(c_ptr) :: landIceMask
     type(c_ptr) :: solveStress
     type(c_ptr) :: solveVelocity
  end type evp_pre_fields

  ! ---- struct evp_category_fields ----
  type, bind(C), public :: evp_category_fields
     integer(c_int) :: nCategories
     type(c_ptr) :: iceAreaCategory
     type(c_ptr) :: iceVolumeCategory
     type(c_ptr) :: snowVolumeCategory
  end type evp_category_fields

  ! ---- struct evp_pre_options ----
  type, bind(C), public :: evp_pre_options
     integer(c_int) :: use_air_stress
     integer(c_int) :: use_surface_tilt
     integer(c_int) :: geostrophic_surface_tilt
     integer(c_int) :: calc_velocity_masks
     integer(c_int) :: cold_start
  end type evp_pre_options

  ! ---- struct evp_post_fields ----
  type, bind(C), public :: evp_post_fields
     type(c_ptr) :: uVelocity
     type(c_ptr) :: vVelocity
     type(c_ptr) :: divergence
     type(c_ptr) :: shear
     type(c_ptr) :: ridgeConvergence
     type(c_ptr) :: ridgeShear
     type(c_ptr) :: principalStress1Var
     type(c_ptr) :: principalStress2Var
     type(c_ptr) :: oceanStressCellU
     type(c_ptr) :: oceanStressCellV
     type(c_ptr) :: oceanStressU
     type(c_ptr) :: oceanStressV
     type(c_ptr) :: oceanStressCoeff
     type(c_ptr) :: principalStress1Weak
     type(c_ptr) :: principalStress2Weak
  end type evp_post_fields

  ! ---- struct evp_weak_mesh ----
  type, bind(C), public :: evp_weak_mesh
     integer(c_int) :: nEdges
     real(c_double) :: sphere_radius
     type(c_ptr) :: edgesOnCell
     type(c_ptr) :: verticesOnEdge
     type(c_ptr) :: edgesOnVertex
     type(c_ptr) :: cellsOnEdge
     type(c_ptr) :: dvEdge
     type(c_ptr) :: dcEdge
     type(c_ptr) :: areaCell
     type(c_ptr) :: areaTriangle
     type(c_ptr) :: normalVectorPolygon
     type(c_ptr) :: normalVectorTriangle
     type(c_ptr) :: latCellRotated
     type(c_ptr) :: latVertexRotated
  end type evp_weak_mesh

  ! ---- struct evp_weak_fields ----
  type, bind(C), public :: evp_weak_fields
     type(c_ptr) :: stress11Weak
     type(c_ptr) :: stress22Weak
     type(c_ptr) :: stress12Weak
     type(c_ptr) :: strain11Weak
     type(c_ptr) :: strain22Weak
     type(c_ptr) :: strain12Weak
     type(c_ptr) :: replacementPressureWeak
  end type evp_weak_fields

  ! ---- functions of include/evp_b200.h ----
  interface

     function evp_create(handle, mesh, options) bind(C, name="evp_create") result(ierr)
       import :: c_ptr, c_int, evp_mesh_desc, evp_options
       type(c_ptr), intent(out) :: handle
       type(evp_mesh_desc), intent(in) :: mesh
       type(evp_options), intent(in) :: options
       integer(c_int) :: ierr
     end function evp_create

     function evp_set_options(handle, options) bind(C, name="evp_set_options") result(ierr)
       import :: c_ptr, c_int, evp_options
       type(c_ptr), value :: handle
       type(evp_options), intent(in) :: options
       integer(c_int) :: ierr
     end function evp_set_options

     function evp_update_step(handle, fields) bind(C, name="evp_update_step") result(ierr)
       import :: c_ptr, c_int, evp_step_fields
       type(c_ptr), value :: handle
       type(evp_step_fields), intent(in) :: fields
       integer(c_int) :: ierr
     end function evp_update_step

     function evp_set_masks(handle, solveStress, solveVelocity) bind(C, name="evp_set_masks") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       type(c_ptr), value :: solveStress
       type(c_ptr), value :: solveVelocity
       integer(c_int) :: ierr
     end function evp_set_masks

     function evp_run_subcycles(handle, nSubcycles) bind(C, name="evp_run_subcycles") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int), value :: nSubcycles
       integer(c_int) :: ierr
     end function evp_run_subcycles

     function evp_synchronize(handle) bind(C, name="evp_synchronize") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int) :: ierr
     end function evp_synchronize

     function evp_fetch(handle, out) bind(C, name="evp_fetch") result(ierr)
       import :: c_ptr, c_int, evp_out_fields
       type(c_ptr), value :: handle
       type(evp_out_fields), intent(in) :: out
       integer(c_int) :: ierr
     end function evp_fetch

     function evp_destroy(handle) bind(C, name="evp_destroy") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int) :: ierr
     end function evp_destroy

     function evp_last_error_string() bind(C, name="evp_last_error_string") result(str)
       import :: c_ptr
       type(c_ptr) :: str
     end function evp_last_error_string

     function evp_comm_get_unique_id(id128) bind(C, name="evp_comm_get_unique_id") result(ierr)
       import :: c_char, c_int
       character(kind=c_char), intent(out) :: id128(128)
       integer(c_int) :: ierr
     end function evp_comm_get_unique_id

     function evp_comm_init(handle, rank, nRanks, id128) bind(C, name="evp_comm_init") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr), value :: handle
       integer(c_int), value :: rank
       integer(c_int), value :: nRanks
       character(kind=c_char), intent(in) :: id128(128)
       integer(c_int) :: ierr
     end function evp_comm_init

     function evp_set_halo(handle, nNeighbours, neighbourRank, sendOffset, sendIndex, recvOffset, recvIndex) &
          bind(C, name="evp_set_halo") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int), value :: nNeighbours
       integer(c_int), intent(in) :: neighbourRank(*)
       integer(c_int), intent(in) :: sendOffset(*)
       integer(c_int), intent(in) :: sendIndex(*)
       integer(c_int), intent(in) :: recvOffset(*)
       integer(c_int), intent(in) :: recvIndex(*)
       integer(c_int) :: ierr
     end function evp_set_halo

     function evp_halo_mode(handle, mode, why, whyLen) bind(C, name="evp_halo_mode") result(ierr)
       import :: c_ptr, c_char, c_int
       type(c_ptr), value :: handle
       integer(c_int), intent(out) :: mode
       character(kind=c_char), intent(out) :: why(*)
       integer(c_int), value :: whyLen
       integer(c_int) :: ierr
     end function evp_halo_mode

     function evp_set_mesh_ext(handle, ext) bind(C, name="evp_set_mesh_ext") result(ierr)
       import :: c_ptr, c_int, evp_mesh_ext
       type(c_ptr), value :: handle
       type(evp_mesh_ext), intent(in) :: ext
       integer(c_int) :: ierr
     end function evp_set_mesh_ext

     function evp_set_state(handle, uVelocity, vVelocity, stress11, stress22, stress12, solveVelocityPrevious) &
          bind(C, name="evp_set_state") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle, uVelocity, vVelocity, stress11, stress22, stress12, solveVelocityPrevious
       integer(c_int) :: ierr
     end function evp_set_state

     function evp_pre_subcycle(handle, fields, options) bind(C, name="evp_pre_subcycle") result(ierr)
       import :: c_ptr, c_int, evp_pre_fields, evp_pre_options
       type(c_ptr), value :: handle
       type(evp_pre_fields), intent(in) :: fields
       type(evp_pre_options), intent(in) :: options
       integer(c_int) :: ierr
     end function evp_pre_subcycle

     function evp_aggregate(handle, categories, hibler_strength) bind(C, name="evp_aggregate") result(ierr)
       import :: c_ptr, c_int, evp_category_fields
       type(c_ptr), value :: handle
       type(evp_category_fields), intent(in) :: categories
       integer(c_int), value :: hibler_strength
       integer(c_int) :: ierr
     end function evp_aggregate

     function evp_fetch_aggregate(handle, iceAreaCell, iceVolumeCell, snowVolumeCell, totalMassCell, icePressure) &
          bind(C, name="evp_fetch_aggregate") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       type(c_ptr), value :: iceAreaCell
       type(c_ptr), value :: iceVolumeCell
       type(c_ptr), value :: snowVolumeCell
       type(c_ptr), value :: totalMassCell
       type(c_ptr), value :: icePressure
       integer(c_int) :: ierr
     end function evp_fetch_aggregate

     function evp_post_subcycle(handle, out) bind(C, name="evp_post_subcycle") result(ierr)
       import :: c_ptr, c_int, evp_post_fields
       type(c_ptr), value :: handle
       type(evp_post_fields), intent(in) :: out
       integer(c_int) :: ierr
     end function evp_post_subcycle

     function evp_set_weak_mesh(handle, mesh) bind(C, name="evp_set_weak_mesh") result(ierr)
       import :: c_ptr, c_int, evp_weak_mesh
       type(c_ptr), value :: handle
       type(evp_weak_mesh), intent(in) :: mesh
       integer(c_int) :: ierr
     end function evp_set_weak_mesh

     function evp_update_weak_state(handle, fields) bind(C, name="evp_update_weak_state") result(ierr)
       import :: c_ptr, c_int, evp_weak_fields
       type(c_ptr), value :: handle
       type(evp_weak_fields), intent(in) :: fields
       integer(c_int) :: ierr
     end function evp_update_weak_state

     function evp_fetch_weak(handle, fields) bind(C, name="evp_fetch_weak") result(ierr)
       import :: c_ptr, c_int, evp_weak_fields
       type(c_ptr), value :: handle
       type(evp_weak_fields), intent(in) :: fields
       integer(c_int) :: ierr
     end function evp_fetch_weak

  end interface

  ! one block per rank (mesh_pool.F:98-105) <-> one handle <-> one GPU
  type(c_ptr) :: evpHandle = c_null_ptr

  integer(c_int) :: nElasticSubcycle = 120

contains

#ifndef EVP_B200_STANDALONE

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_supported
!
!> \brief Is this namelist combination covered by the CUDA path?  (velocity_solver.F:168-198)
!-----------------------------------------------------------------------

  function seaice_evp_b200_supported(domain) result(supported)

    type(domain_type), intent(in) :: domain
    logical :: supported

    character(len=strKIND), pointer :: &
         config_strain_scheme, &
         config_stress_divergence_scheme

    call MPAS_pool_get_config(domain % configs, "config_strain_scheme", config_strain_scheme)
    call MPAS_pool_get_config(domain % configs, "config_stress_divergence_scheme", config_stress_divergence_scheme)

    ! variational / variational (with or without config_average_variational_strain, which needs
    ! seaice_evp_b200_set_mesh_ext for areaCell), weak / weak and weak strain + variational divergence (both need
    ! seaice_evp_b200_set_weak_mesh) are all covered; in a pure weak run pkgVariational is inactive and the
    ! velocity_variational fields are simply not passed (c_null_ptr)
    supported = trim(config_stress_divergence_scheme) == "variational" .or. &
                trim(config_strain_scheme) == "weak"

  end function seaice_evp_b200_supported

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  evp_b200_check
!
!> \brief C status -> MPAS critical error, the reference's convention (velocity_solver.F:262-265)
!-----------------------------------------------------------------------

  subroutine evp_b200_check(ierr, where)

    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where

    character(kind=c_char), dimension(:), pointer :: cmsg
    character(len=512) :: msg
    integer :: i

    if (ierr /= EVP_OK) then
       msg = " "
       call c_f_pointer(evp_last_error_string(), cmsg, [512])
       do i = 1, 512
          if (cmsg(i) == c_null_char) exit
          msg(i:i) = cmsg(i)
       enddo
       call mpas_log_write("libevp_b200: "//trim(where)//": "//trim(msg), MPAS_LOG_CRIT)
    endif

  end subroutine evp_b200_check

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  fill_options
!-----------------------------------------------------------------------

  subroutine fill_options(domain, options)

    use seaice_velocity_solver_constitutive_relation, only: &
         constitutiveRelationType, dampingTimescale, numericalInertiaCoefficient

    type(domain_type), intent(in) :: domain
    type(evp_options), intent(out) :: options

    type(MPAS_pool_type), pointer :: velocitySolverPool
    real(kind=RKIND), pointer :: elasticTimeStep, dynamicsTimeStep
    character(len=strKIND), pointer :: config_ocean_stress_type, config_strain_scheme, config_stress_divergence_scheme
    logical, pointer :: config_use_ocean_stress, config_use_special_boundaries_velocity, &
         config_average_variational_strain
    integer, pointer :: config_elastic_subcycle_number

    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_solver", velocitySolverPool)
    call MPAS_pool_get_array(velocitySolverPool, "elasticTimeStep", elasticTimeStep)
    call MPAS_pool_get_array(velocitySolverPool, "dynamicsTimeStep", dynamicsTimeStep)
    call MPAS_pool_get_config(domain % configs, "config_ocean_stress_type", config_ocean_stress_type)
    call MPAS_pool_get_config(domain % configs, "config_use_ocean_stress", config_use_ocean_stress)
    call MPAS_pool_get_config(domain % configs, "config_use_special_boundaries_velocity", &
                                                 config_use_special_boundaries_velocity)
    call MPAS_pool_get_config(domain % configs, "config_elastic_subcycle_number", config_elastic_subcycle_number)
    call MPAS_pool_get_config(domain % configs, "config_average_variational_strain", config_average_variational_strain)
    call MPAS_pool_get_config(domain % configs, "config_strain_scheme", config_strain_scheme)
    call MPAS_pool_get_config(domain % configs, "config_stress_divergence_scheme", config_stress_divergence_scheme)

    nElasticSubcycle = config_elastic_subcycle_number

    ! EVP_CR_* are numbered like the reference's own constants (constitutive_relation.F:34-38)
    options % constitutive_relation_type = constitutiveRelationType
    if (trim(config_ocean_stress_type) == "quadratic") then
       options % ocean_stress_type = EVP_OCEAN_QUADRATIC
    else
       options % ocean_stress_type = EVP_OCEAN_LINEAR
    endif
    options % use_ocean_stress = merge(1, 0, config_use_ocean_stress)
    options % use_special_boundaries_velocity = merge(1, 0, config_use_special_boundaries_velocity)
    options % device = -1                    ! the device the host selected (cudaSetDevice / CUDA_VISIBLE_DEVICES)
    options % flags = EVP_FLAG_PIN_HOST      ! MPAS pool arrays live at stable addresses
    options % average_variational_strain = merge(1, 0, config_average_variational_strain)
    options % strain_scheme = merge(EVP_SCHEME_WEAK, EVP_SCHEME_VARIATIONAL, trim(config_strain_scheme) == "weak")
    options % stress_divergence_scheme = merge(EVP_SCHEME_WEAK, EVP_SCHEME_VARIATIONAL, &
                                               trim(config_stress_divergence_scheme) == "weak")
    options % elasticTimeStep = elasticTimeStep
    options % dynamicsTimeStep = dynamicsTimeStep
    options % dampingTimescale = dampingTimescale
    options % numericalInertiaCoefficient = numericalInertiaCoefficient

  end subroutine fill_options

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_create
!
!> \brief seaice_mesh_pool_create equivalent: upload connectivity + basis, build the handle
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_create(domain)

    type(domain_type), intent(inout) :: domain

    type(MPAS_pool_type), pointer :: meshPool, velocityVariationalPool, specialBoundariesPool
    type(evp_mesh_desc) :: mesh
    type(evp_options) :: options

    integer, pointer :: nCells, nCellsSolve, nVertices, nVerticesSolve, maxEdges, vertexDegree
    integer, dimension(:), pointer :: nEdgesOnCell, vertexBoundaryType, vertexBoundarySourceLocal
    integer, dimension(:,:), pointer :: verticesOnCell, cellsOnVertex, cellVerticesAtVertex
    real(kind=RKIND), dimension(:), pointer :: tanLatVertexRotatedOverRadius, variationalDenominator
    real(kind=RKIND), dimension(:,:,:), pointer :: &
         basisGradientU, basisGradientV, basisIntegralsU, basisIntegralsV, basisIntegralsMetric
    logical, pointer :: config_use_special_boundaries_velocity, pkgVariationalActive

    call MPAS_pool_get_subpool(domain % blocklist % structs, "mesh", meshPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_variational", velocityVariationalPool)

    call MPAS_pool_get_dimension(meshPool, "nCells", nCells)
    call MPAS_pool_get_dimension(meshPool, "nCellsSolve", nCellsSolve)
    call MPAS_pool_get_dimension(meshPool, "nVertices", nVertices)
    call MPAS_pool_get_dimension(meshPool, "nVerticesSolve", nVerticesSolve)
    call MPAS_pool_get_dimension(meshPool, "maxEdges", maxEdges)
    call MPAS_pool_get_dimension(meshPool, "vertexDegree", vertexDegree)

    call MPAS_pool_get_array(meshPool, "nEdgesOnCell", nEdgesOnCell)
    call MPAS_pool_get_array(meshPool, "verticesOnCell", verticesOnCell)
    call MPAS_pool_get_array(meshPool, "cellsOnVertex", cellsOnVertex)

    ! a pure weak configuration has no velocity_variational fields (pkgVariational inactive,
    ! src/model_forward/mpas_seaice_core_interface.F:158-185): every pointer below stays c_null_ptr
    mesh % cellVerticesAtVertex = c_null_ptr
    mesh % basisGradientU = c_null_ptr
    mesh % basisGradientV = c_null_ptr
    mesh % basisIntegralsU = c_null_ptr
    mesh % basisIntegralsV = c_null_ptr
    mesh % basisIntegralsMetric = c_null_ptr
    mesh % tanLatVertexRotatedOverRadius = c_null_ptr
    mesh % variationalDenominator = c_null_ptr
    call MPAS_pool_get_package(domain % packages, "pkgVariationalActive", pkgVariationalActive)
    if (pkgVariationalActive) then
    call MPAS_pool_get_array(velocityVariationalPool, "cellVerticesAtVertex", cellVerticesAtVertex)
    call MPAS_pool_get_array(velocityVariationalPool, "basisGradientU", basisGradientU)
    call MPAS_pool_get_array(velocityVariationalPool, "basisGradientV", basisGradientV)
    call MPAS_pool_get_array(velocityVariationalPool, "basisIntegralsU", basisIntegralsU)
    call MPAS_pool_get_array(velocityVariationalPool, "basisIntegralsV", basisIntegralsV)
    call MPAS_pool_get_array(velocityVariationalPool, "basisIntegralsMetric", basisIntegralsMetric)
    call MPAS_pool_get_array(velocityVariationalPool, "tanLatVertexRotatedOverRadius", tanLatVertexRotatedOverRadius)
    ! filled by variational_denominator (variational.F:358-445); read the same way at velocity_solver.F:2718
    call MPAS_pool_get_array(velocityVariationalPool, "variationalDenominator", variationalDenominator)
    mesh % cellVerticesAtVertex = c_loc(cellVerticesAtVertex)
    mesh % basisGradientU = c_loc(basisGradientU)
    mesh % basisGradientV = c_loc(basisGradientV)
    mesh % basisIntegralsU = c_loc(basisIntegralsU)
    mesh % basisIntegralsV = c_loc(basisIntegralsV)
    mesh % basisIntegralsMetric = c_loc(basisIntegralsMetric)
    mesh % tanLatVertexRotatedOverRadius = c_loc(tanLatVertexRotatedOverRadius)
    mesh % variationalDenominator = c_loc(variationalDenominator)
    endif

    mesh % nCells = nCells
    mesh % nCellsSolve = nCellsSolve
    mesh % nVertices = nVertices
    mesh % nVerticesSolve = nVerticesSolve
    mesh % maxEdges = maxEdges
    mesh % vertexDegree = vertexDegree
    mesh % nEdgesOnCell = c_loc(nEdgesOnCell)
    mesh % verticesOnCell = c_loc(verticesOnCell)
    mesh % cellsOnVertex = c_loc(cellsOnVertex)
    mesh % vertexBoundaryType = c_null_ptr
    mesh % vertexBoundarySourceLocal = c_null_ptr

    call MPAS_pool_get_config(domain % configs, "config_use_special_boundaries_velocity", &
                                                 config_use_special_boundaries_velocity)
    if (config_use_special_boundaries_velocity) then
       call MPAS_pool_get_subpool(domain % blocklist % structs, "special_boundaries", specialBoundariesPool)
       call MPAS_pool_get_array(specialBoundariesPool, "vertexBoundaryType", vertexBoundaryType)
       call MPAS_pool_get_array(specialBoundariesPool, "vertexBoundarySourceLocal", vertexBoundarySourceLocal)
       mesh % vertexBoundaryType = c_loc(vertexBoundaryType)
       mesh % vertexBoundarySourceLocal = c_loc(vertexBoundarySourceLocal)
    endif

    call fill_options(domain, options)

    call evp_b200_check(evp_create(evpHandle, mesh, options), "evp_create")

    ! multi-rank runs: the velocityHaloExchangeGroup lists (velocity_solver.F:259-349) go to
    ! evp_set_halo here; see INTEGRATION.md section 4 for how they are read off the dmpar exchange lists.

  end subroutine seaice_evp_b200_create

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_update
!
!> \brief seaice_mesh_pool_update equivalent: one dynamics step's inputs, host -> device
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_update(domain)

    type(domain_type), intent(inout) :: domain

    type(MPAS_pool_type), pointer :: velocitySolverPool, velocityVariationalPool, icestatePool
    type(evp_step_fields) :: f
    type(evp_options) :: options
    type(evp_weak_fields) :: wf

    integer, dimension(:), pointer :: solveStress, solveVelocity
    real(kind=RKIND), dimension(:), pointer :: &
         icePressure, uVelocity, vVelocity, totalMassVertex, totalMassVertexfVertex, iceAreaVertex, &
         airStressVertexU, airStressVertexV, surfaceTiltForceU, surfaceTiltForceV, oceanStressU, oceanStressV, &
         uOceanVelocityVertex, vOceanVelocityVertex, uVelocityInitial, vVelocityInitial
    real(kind=RKIND), dimension(:,:), pointer :: stress11, stress22, stress12
    logical, pointer :: pkgVariationalActive

    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_solver", velocitySolverPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_variational", velocityVariationalPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "icestate", icestatePool)

    call MPAS_pool_get_array(velocitySolverPool, "solveStress", solveStress)
    call MPAS_pool_get_array(velocitySolverPool, "solveVelocity", solveVelocity)
    call MPAS_pool_get_array(velocitySolverPool, "icePressure", icePressure)
    call MPAS_pool_get_array(velocitySolverPool, "uVelocity", uVelocity)
    call MPAS_pool_get_array(velocitySolverPool, "vVelocity", vVelocity)
    call MPAS_pool_get_package(domain % packages, "pkgVariationalActive", pkgVariationalActive)
    if (pkgVariationalActive) then
       call MPAS_pool_get_array(velocityVariationalPool, "stress11", stress11)
       call MPAS_pool_get_array(velocityVariationalPool, "stress22", stress22)
       call MPAS_pool_get_array(velocityVariationalPool, "stress12", stress12)
    endif
    call MPAS_pool_get_array(icestatePool, "totalMassVertex", totalMassVertex)
    call MPAS_pool_get_array(icestatePool, "iceAreaVertex", iceAreaVertex)
    call MPAS_pool_get_array(velocitySolverPool, "totalMassVertexfVertex", totalMassVertexfVertex)
    call MPAS_pool_get_array(velocitySolverPool, "airStressVertexU", airStressVertexU)
    call MPAS_pool_get_array(velocitySolverPool, "airStressVertexV", airStressVertexV)
    call MPAS_pool_get_array(velocitySolverPool, "surfaceTiltForceU", surfaceTiltForceU)
    call MPAS_pool_get_array(velocitySolverPool, "surfaceTiltForceV", surfaceTiltForceV)
    call MPAS_pool_get_array(velocitySolverPool, "oceanStressU", oceanStressU)
    call MPAS_pool_get_array(velocitySolverPool, "oceanStressV", oceanStressV)
    call MPAS_pool_get_array(velocitySolverPool, "uOceanVelocityVertex", uOceanVelocityVertex)
    call MPAS_pool_get_array(velocitySolverPool, "vOceanVelocityVertex", vOceanVelocityVertex)
    call MPAS_pool_get_array(velocitySolverPool, "uVelocityInitial", uVelocityInitial)
    call MPAS_pool_get_array(velocitySolverPool, "vVelocityInitial", vVelocityInitial)

    f % solveStress = c_loc(solveStress)
    f % solveVelocity = c_loc(solveVelocity)
    f % icePressure = c_loc(icePressure)
    f % uVelocity = c_loc(uVelocity)
    f % vVelocity = c_loc(vVelocity)
    f % stress11 = c_null_ptr
    f % stress22 = c_null_ptr
    f % stress12 = c_null_ptr
    if (pkgVariationalActive) then
       f % stress11 = c_loc(stress11)
       f % stress22 = c_loc(stress22)
       f % stress12 = c_loc(stress12)
    endif
    f % totalMassVertex = c_loc(totalMassVertex)
    f % totalMassVertexfVertex = c_loc(totalMassVertexfVertex)
    f % iceAreaVertex = c_loc(iceAreaVertex)
    f % airStressVertexU = c_loc(airStressVertexU)
    f % airStressVertexV = c_loc(airStressVertexV)
    f % surfaceTiltForceU = c_loc(surfaceTiltForceU)
    f % surfaceTiltForceV = c_loc(surfaceTiltForceV)
    f % oceanStressU = c_loc(oceanStressU)
    f % oceanStressV = c_loc(oceanStressV)
    f % uOceanVelocityVertex = c_loc(uOceanVelocityVertex)
    f % vOceanVelocityVertex = c_loc(vOceanVelocityVertex)
    f % uVelocityInitial = c_loc(uVelocityInitial)
    f % vVelocityInitial = c_loc(vVelocityInitial)

    ! config_dt may change between steps (coupled runs): refresh the scalars, cheap
    call fill_options(domain, options)
    call evp_b200_check(evp_set_options(evpHandle, options), "evp_set_options")

    call evp_b200_check(evp_update_step(evpHandle, f), "evp_update_step")

    if (options % stress_divergence_scheme == EVP_SCHEME_WEAK) then
       call weak_state_pointers(domain, wf)
       call evp_b200_check(evp_update_weak_state(evpHandle, wf), "evp_update_weak_state")
    endif

  end subroutine seaice_evp_b200_update

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_subcycle
!
!> \brief Replacement body of subcycle_velocity_solver (velocity_solver.F:2404-2464)
!>
!> The special-boundary velocity copies before the loop and after every subcycle, the
!> config_elastic_subcycle_number subcycles and the per-subcycle halo exchange all run inside one
!> CUDA graph on the device; afterwards everything velocity_solver_post_subcycle and the restart
!> stream read is copied back into the pool arrays.
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_subcycle(domain)

    use seaice_special_boundaries, only: &
         seaice_set_special_boundaries_velocity_masks

    type(domain_type), intent(inout) :: domain

    type(MPAS_pool_type), pointer :: velocitySolverPool, velocityVariationalPool
    type(evp_out_fields) :: o
    type(evp_weak_fields) :: wf
    character(len=strKIND), pointer :: config_strain_scheme

    integer, dimension(:), pointer :: solveStress, solveVelocity
    real(kind=RKIND), dimension(:), pointer :: &
         uVelocity, vVelocity, stressDivergenceU, stressDivergenceV, oceanStressCoeff
    real(kind=RKIND), dimension(:,:), pointer :: &
         stress11, stress22, stress12, strain11, strain22, strain12, replacementPressure
    logical, pointer :: config_use_special_boundaries_velocity_masks, pkgVariationalActive

    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_solver", velocitySolverPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_variational", velocityVariationalPool)

    ! the mask variant of the special boundaries only overwrites the two integer masks with constant
    ! arrays (special_boundaries.F:345-401): apply it once on the host, then push the masks
    call MPAS_pool_get_config(domain % configs, "config_use_special_boundaries_velocity_masks", &
                                                 config_use_special_boundaries_velocity_masks)
    if (config_use_special_boundaries_velocity_masks) then
       call seaice_set_special_boundaries_velocity_masks(domain)
       call MPAS_pool_get_array(velocitySolverPool, "solveStress", solveStress)
       call MPAS_pool_get_array(velocitySolverPool, "solveVelocity", solveVelocity)
       call evp_b200_check(evp_set_masks(evpHandle, c_loc(solveStress), c_loc(solveVelocity)), "evp_set_masks")
    endif

    call evp_b200_check(evp_run_subcycles(evpHandle, nElasticSubcycle), "evp_run_subcycles")

    call MPAS_pool_get_array(velocitySolverPool, "uVelocity", uVelocity)
    call MPAS_pool_get_array(velocitySolverPool, "vVelocity", vVelocity)
    call MPAS_pool_get_array(velocitySolverPool, "stressDivergenceU", stressDivergenceU)
    call MPAS_pool_get_array(velocitySolverPool, "stressDivergenceV", stressDivergenceV)
    call MPAS_pool_get_array(velocitySolverPool, "oceanStressCoeff", oceanStressCoeff)
    call MPAS_pool_get_package(domain % packages, "pkgVariationalActive", pkgVariationalActive)
    if (pkgVariationalActive) then
       call MPAS_pool_get_array(velocityVariationalPool, "stress11", stress11)
       call MPAS_pool_get_array(velocityVariationalPool, "stress22", stress22)
       call MPAS_pool_get_array(velocityVariationalPool, "stress12", stress12)
       call MPAS_pool_get_array(velocityVariationalPool, "strain11", strain11)
       call MPAS_pool_get_array(velocityVariationalPool, "strain22", strain22)
       call MPAS_pool_get_array(velocityVariationalPool, "strain12", strain12)
       call MPAS_pool_get_array(velocityVariationalPool, "replacementPressure", replacementPressure)
    endif

    o % uVelocity = c_loc(uVelocity)
    o % vVelocity = c_loc(vVelocity)
    o % stress11 = c_null_ptr
    o % stress22 = c_null_ptr
    o % stress12 = c_null_ptr
    o % strain11 = c_null_ptr
    o % strain22 = c_null_ptr
    o % strain12 = c_null_ptr
    o % replacementPressure = c_null_ptr
    if (pkgVariationalActive) then
       o % stress11 = c_loc(stress11)
       o % stress22 = c_loc(stress22)
       o % stress12 = c_loc(stress12)
       o % strain11 = c_loc(strain11)
       o % strain22 = c_loc(strain22)
       o % strain12 = c_loc(strain12)
       o % replacementPressure = c_loc(replacementPressure)
    endif
    o % stressDivergenceU = c_loc(stressDivergenceU)
    o % stressDivergenceV = c_loc(stressDivergenceV)
    o % oceanStressCoeff = c_loc(oceanStressCoeff)

    call evp_b200_check(evp_fetch(evpHandle, o), "evp_fetch")

    call MPAS_pool_get_config(domain % configs, "config_strain_scheme", config_strain_scheme)
    if (trim(config_strain_scheme) == "weak") then
       call weak_state_pointers(domain, wf)
       call evp_b200_check(evp_fetch_weak(evpHandle, wf), "evp_fetch_weak")
    endif

  end subroutine seaice_evp_b200_subcycle

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_set_mesh_ext
!
!> \brief Mesh fields velocity_solver_pre/post_subcycle read (once, after seaice_evp_b200_create)
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_set_mesh_ext(domain)

    type(domain_type), intent(inout) :: domain

    type(MPAS_pool_type), pointer :: meshPool, boundaryPool, oceanCouplingPool
    type(evp_mesh_ext) :: ext
    integer, dimension(:,:), pointer :: cellsOnCell
    integer, dimension(:), pointer :: interiorVertex, landIceMaskVertex
    real(kind=RKIND), dimension(:), pointer :: areaCell, areaTriangle, fVertex

    call MPAS_pool_get_subpool(domain % blocklist % structs, "mesh", meshPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "boundary", boundaryPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "ocean_coupling", oceanCouplingPool)

    call MPAS_pool_get_array(meshPool, "cellsOnCell", cellsOnCell)
    call MPAS_pool_get_array(meshPool, "areaCell", areaCell)
    call MPAS_pool_get_array(meshPool, "areaTriangle", areaTriangle)
    call MPAS_pool_get_array(meshPool, "fVertex", fVertex)
    call MPAS_pool_get_array(boundaryPool, "interiorVertex", interiorVertex)
    call MPAS_pool_get_array(oceanCouplingPool, "landIceMaskVertex", landIceMaskVertex)

    ext % cellsOnCell = c_loc(cellsOnCell)
    ext % interiorVertex = c_loc(interiorVertex)
    ext % landIceMaskVertex = c_loc(landIceMaskVertex)
    ext % areaCell = c_loc(areaCell)
    ext % areaTriangle = c_loc(areaTriangle)
    ext % fVertex = c_loc(fVertex)

    call evp_b200_check(evp_set_mesh_ext(evpHandle, ext), "evp_set_mesh_ext")

  end subroutine seaice_evp_b200_set_mesh_ext

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_set_weak_mesh
!
!> \brief Edge connectivity + normal vectors of the weak operators (once, after seaice_evp_b200_create and
!>        after seaice_init_velocity_solver_weak filled normalVectorPolygon / normalVectorTriangle)
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_set_weak_mesh(domain)

    type(domain_type), intent(inout) :: domain

    type(MPAS_pool_type), pointer :: meshPool, velocityWeakPool
    type(evp_weak_mesh) :: wm
    integer, pointer :: nEdges
    real(kind=RKIND), pointer :: sphere_radius
    integer, dimension(:,:), pointer :: edgesOnCell, verticesOnEdge, edgesOnVertex, cellsOnEdge
    real(kind=RKIND), dimension(:), pointer :: dvEdge, dcEdge, areaCell, areaTriangle, latCellRotated, latVertexRotated
    real(kind=RKIND), dimension(:,:,:), pointer :: normalVectorPolygon, normalVectorTriangle

    call MPAS_pool_get_subpool(domain % blocklist % structs, "mesh", meshPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_weak", velocityWeakPool)

    call MPAS_pool_get_dimension(meshPool, "nEdges", nEdges)
    call MPAS_pool_get_config(meshPool, "sphere_radius", sphere_radius)
    call MPAS_pool_get_array(meshPool, "edgesOnCell", edgesOnCell)
    call MPAS_pool_get_array(meshPool, "verticesOnEdge", verticesOnEdge)
    call MPAS_pool_get_array(meshPool, "edgesOnVertex", edgesOnVertex)
    call MPAS_pool_get_array(meshPool, "cellsOnEdge", cellsOnEdge)
    call MPAS_pool_get_array(meshPool, "dvEdge", dvEdge)
    call MPAS_pool_get_array(meshPool, "dcEdge", dcEdge)
    call MPAS_pool_get_array(meshPool, "areaCell", areaCell)
    call MPAS_pool_get_array(meshPool, "areaTriangle", areaTriangle)
    call MPAS_pool_get_array(velocityWeakPool, "normalVectorPolygon", normalVectorPolygon)
    call MPAS_pool_get_array(velocityWeakPool, "normalVectorTriangle", normalVectorTriangle)
    call MPAS_pool_get_array(velocityWeakPool, "latCellRotated", latCellRotated)
    call MPAS_pool_get_array(velocityWeakPool, "latVertexRotated", latVertexRotated)

    wm % nEdges = nEdges
    wm % sphere_radius = sphere_radius
    wm % edgesOnCell = c_loc(edgesOnCell)
    wm % verticesOnEdge = c_loc(verticesOnEdge)
    wm % edgesOnVertex = c_loc(edgesOnVertex)
    wm % cellsOnEdge = c_loc(cellsOnEdge)
    wm % dvEdge = c_loc(dvEdge)
    wm % dcEdge = c_loc(dcEdge)
    wm % areaCell = c_loc(areaCell)
    wm % areaTriangle = c_loc(areaTriangle)
    wm % normalVectorPolygon = c_loc(normalVectorPolygon)
    wm % normalVectorTriangle = c_loc(normalVectorTriangle)
    wm % latCellRotated = c_loc(latCellRotated)
    wm % latVertexRotated = c_loc(latVertexRotated)

    call evp_b200_check(evp_set_weak_mesh(evpHandle, wm), "evp_set_weak_mesh")

  end subroutine seaice_evp_b200_set_weak_mesh

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  weak_state_pointers
!
!> \brief The velocity_weak pool fields as evp_weak_fields (src/Registry.xml:3730-3745)
!-----------------------------------------------------------------------

  subroutine weak_state_pointers(domain, wf)

    type(domain_type), intent(inout) :: domain
    type(evp_weak_fields), intent(out) :: wf

    type(MPAS_pool_type), pointer :: velocityWeakPool
    real(kind=RKIND), dimension(:), pointer :: stress11, stress22, stress12, strain11, strain22, strain12, &
         replacementPressure

    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_weak", velocityWeakPool)
    call MPAS_pool_get_array(velocityWeakPool, "stress11", stress11)
    call MPAS_pool_get_array(velocityWeakPool, "stress22", stress22)
    call MPAS_pool_get_array(velocityWeakPool, "stress12", stress12)
    call MPAS_pool_get_array(velocityWeakPool, "strain11", strain11)
    call MPAS_pool_get_array(velocityWeakPool, "strain22", strain22)
    call MPAS_pool_get_array(velocityWeakPool, "strain12", strain12)
    call MPAS_pool_get_array(velocityWeakPool, "replacementPressure", replacementPressure)

    wf % stress11Weak = c_loc(stress11)
    wf % stress22Weak = c_loc(stress22)
    wf % stress12Weak = c_loc(stress12)
    wf % strain11Weak = c_loc(strain11)
    wf % strain22Weak = c_loc(strain22)
    wf % strain12Weak = c_loc(strain12)
    wf % replacementPressureWeak = c_loc(replacementPressure)

  end subroutine weak_state_pointers

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_step
!
!> \brief One dynamics step with pre-subcycle, subcycle and post-subcycle on the device
!>
!> Replaces, inside seaice_run_velocity_solver (velocity_solver.F:577-593), everything after
!> aggregate_mass_and_area and the (unmasked) ice strength -- see INTEGRATION.md section 6.  The caller
!> has filled iceAreaCell / iceAreaCellInitial / totalMassCell (aggregate_mass_and_area, :685-752 and
!> :820-836) and icePressure for every cell that may be solved (ice_strength, :1419-1460, evaluated
!> without its solveStress test; the device applies the mask).  u, v, the stresses and
!> solveVelocityPrevious stay on the device between steps; coldStart = .true. on the first step of a
!> run without a restart file (solveVelocityPrevious = 0 as in the reference, so solved vertices start at
!> the ocean velocity, velocity_solver.F:1252-1258), after a restart call evp_set_state first.
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_step(domain, coldStart)

    type(domain_type), intent(inout) :: domain
    logical, intent(in) :: coldStart

    type(MPAS_pool_type), pointer :: velocitySolverPool, icestatePool, tracersAggregatePool, &
         oceanCouplingPool, atmosCouplingPool, ridgingPool
    type(evp_pre_fields) :: f
    type(evp_pre_options) :: po
    type(evp_post_fields) :: o
    type(evp_options) :: options

    real(kind=RKIND), dimension(:), pointer :: &
         iceAreaCellInitial, iceAreaCell, totalMassCell, icePressure, uOceanVelocity, vOceanVelocity, &
         airStressCellU, airStressCellV, uAirVelocity, vAirVelocity, airDensity, seaSurfaceTiltU, seaSurfaceTiltV, &
         uVelocity, vVelocity, divergence, shear, ridgeConvergence, ridgeShear, oceanStressCellU, oceanStressCellV
    integer, dimension(:), pointer :: landIceMask, solveStress, solveVelocity
    logical, pointer :: config_use_air_stress, config_use_surface_tilt, config_geostrophic_surface_tilt, &
         config_calc_velocity_masks, config_use_column_package, config_use_column_vertical_thermodynamics

    call MPAS_pool_get_config(domain % configs, "config_use_air_stress", config_use_air_stress)
    call MPAS_pool_get_config(domain % configs, "config_use_surface_tilt", config_use_surface_tilt)
    call MPAS_pool_get_config(domain % configs, "config_geostrophic_surface_tilt", config_geostrophic_surface_tilt)
    call MPAS_pool_get_config(domain % configs, "config_calc_velocity_masks", config_calc_velocity_masks)
    call MPAS_pool_get_config(domain % configs, "config_use_column_package", config_use_column_package)
    call MPAS_pool_get_config(domain % configs, "config_use_column_vertical_thermodynamics", &
                                                 config_use_column_vertical_thermodynamics)

    call MPAS_pool_get_subpool(domain % blocklist % structs, "velocity_solver", velocitySolverPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "icestate", icestatePool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "tracers_aggregate", tracersAggregatePool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "ocean_coupling", oceanCouplingPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "atmos_coupling", atmosCouplingPool)
    call MPAS_pool_get_subpool(domain % blocklist % structs, "ridging", ridgingPool)

    call MPAS_pool_get_array(icestatePool, "iceAreaCellInitial", iceAreaCellInitial)
    call MPAS_pool_get_array(tracersAggregatePool, "iceAreaCell", iceAreaCell)
    call MPAS_pool_get_array(icestatePool, "totalMassCell", totalMassCell)
    call MPAS_pool_get_array(velocitySolverPool, "icePressure", icePressure)
    call MPAS_pool_get_array(oceanCouplingPool, "uOceanVelocity", uOceanVelocity)
    call MPAS_pool_get_array(oceanCouplingPool, "vOceanVelocity", vOceanVelocity)
    call MPAS_pool_get_array(oceanCouplingPool, "seaSurfaceTiltU", seaSurfaceTiltU)
    call MPAS_pool_get_array(oceanCouplingPool, "seaSurfaceTiltV", seaSurfaceTiltV)
    call MPAS_pool_get_array(oceanCouplingPool, "landIceMask", landIceMask)
    call MPAS_pool_get_array(velocitySolverPool, "airStressCellU", airStressCellU)
    call MPAS_pool_get_array(velocitySolverPool, "airStressCellV", airStressCellV)
    call MPAS_pool_get_array(atmosCouplingPool, "uAirVelocity", uAirVelocity)
    call MPAS_pool_get_array(atmosCouplingPool, "vAirVelocity", vAirVelocity)
    call MPAS_pool_get_array(atmosCouplingPool, "airDensity", airDensity)
    call MPAS_pool_get_array(velocitySolverPool, "solveStress", solveStress)
    call MPAS_pool_get_array(velocitySolverPool, "solveVelocity", solveVelocity)

    f % iceAreaCellInitial = c_loc(iceAreaCellInitial)
    f % iceAreaCell = c_loc(iceAreaCell)
    f % totalMassCell = c_loc(totalMassCell)
    f % icePressure = c_loc(icePressure)
    f % uOceanVelocity = c_loc(uOceanVelocity)
    f % vOceanVelocity = c_loc(vOceanVelocity)
    ! air_stress (velocity_solver.F:1560-1580): constant_air_stress without the column thermodynamics,
    ! else the stresses the column package / coupler left in airStressCellU/V
    if (.not. config_use_column_package .or. &
         (config_use_column_package .and. .not. config_use_column_vertical_thermodynamics)) then
       f % airStressCellU = c_null_ptr
       f % airStressCellV = c_null_ptr
       f % uAirVelocity = c_loc(uAirVelocity)
       f % vAirVelocity = c_loc(vAirVelocity)
       f % airDensity = c_loc(airDensity)
    else
       f % airStressCellU = c_loc(airStressCellU)
       f % airStressCellV = c_loc(airStressCellV)
       f % uAirVelocity = c_null_ptr
       f % vAirVelocity = c_null_ptr
       f % airDensity = c_null_ptr
    endif
    f % seaSurfaceTiltU = c_loc(seaSurfaceTiltU)
    f % seaSurfaceTiltV = c_loc(seaSurfaceTiltV)
    f % landIceMask = c_loc(landIceMask)
    f % solveStress = c_loc(solveStress)          ! read only when config_calc_velocity_masks = .false.
    f % solveVelocity = c_loc(solveVelocity)

    po % use_air_stress = merge(1, 0, config_use_air_stress)
    po % use_surface_tilt = merge(1, 0, config_use_surface_tilt)
    po % geostrophic_surface_tilt = merge(1, 0, config_geostrophic_surface_tilt)
    po % calc_velocity_masks = merge(1, 0, config_calc_velocity_masks)
    po % cold_start = merge(EVP_START_FIRST_STEP, EVP_START_RESIDENT, coldStart)

    call fill_options(domain, options)
    call evp_b200_check(evp_set_options(evpHandle, options), "evp_set_options")
    call evp_b200_check(evp_pre_subcycle(evpHandle, f, po), "evp_pre_subcycle")
    call evp_b200_check(evp_run_subcycles(evpHandle, nElasticSubcycle), "evp_run_subcycles")

    ! what the rest of the model reads every step: advection (u, v), ridging (divergence, shear,
    ! ridgeConvergence, ridgeShear), coupler (oceanStressCellU/V).  Restart / output fields are fetched
    ! with evp_fetch / evp_post_subcycle only when those streams fire.
    call MPAS_pool_get_array(velocitySolverPool, "uVelocity", uVelocity)
    call MPAS_pool_get_array(velocitySolverPool, "vVelocity", vVelocity)
    call MPAS_pool_get_array(velocitySolverPool, "divergence", divergence)
    call MPAS_pool_get_array(velocitySolverPool, "shear", shear)
    call MPAS_pool_get_array(ridgingPool, "ridgeConvergence", ridgeConvergence)
    call MPAS_pool_get_array(ridgingPool, "ridgeShear", ridgeShear)
    call MPAS_pool_get_array(velocitySolverPool, "oceanStressCellU", oceanStressCellU)
    call MPAS_pool_get_array(velocitySolverPool, "oceanStressCellV", oceanStressCellV)

    o % uVelocity = c_loc(uVelocity)
    o % vVelocity = c_loc(vVelocity)
    o % divergence = c_loc(divergence)
    o % shear = c_loc(shear)
    o % ridgeConvergence = c_loc(ridgeConvergence)
    o % ridgeShear = c_loc(ridgeShear)
    o % principalStress1Var = c_null_ptr
    o % principalStress2Var = c_null_ptr
    o % oceanStressCellU = c_loc(oceanStressCellU)
    o % oceanStressCellV = c_loc(oceanStressCellV)
    o % oceanStressU = c_null_ptr
    o % oceanStressV = c_null_ptr
    o % oceanStressCoeff = c_null_ptr
    o % principalStress1Weak = c_null_ptr
    o % principalStress2Weak = c_null_ptr

    call evp_b200_check(evp_post_subcycle(evpHandle, o), "evp_post_subcycle")

  end subroutine seaice_evp_b200_step

!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!  seaice_evp_b200_destroy
!
!> \brief seaice_mesh_pool_destroy equivalent
!-----------------------------------------------------------------------

  subroutine seaice_evp_b200_destroy(err)

    integer, intent(out) :: err

    err = evp_destroy(evpHandle)
    evpHandle = c_null_ptr

  end subroutine seaice_evp_b200_destroy

#endif

end module seaice_evp_b200
