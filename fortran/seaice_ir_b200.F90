!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!
!  seaice_ir_b200
!
!> \brief  ISO_C_BINDING shim: incremental-remapping transport on a B200 (include/ir_b200.h)
!> \details
!>  What a maintainer adds to src/shared to run incremental_remap_block on the device.  The module is a
!>  sibling of seaice_advection_incremental_remap: it reads the same pools (mesh, incremental_remap,
!>  velocity_solver) and walks the same ordered tracer linked list (tracersHead of
!>  seaice_advection_incremental_remap_tracers), flattens the list into the ir_tracer_desc table of
!>  the C ABI, and replaces the body of the block loop of seaice_run_advection_incremental_remap
!>  (mpas_seaice_advection_incremental_remap.F:2518-2533):
!>
!>     call incremental_remap_block(domain, block, dynamicsTimeStep, tracersHead)
!>  becomes
!>     call seaice_ir_b200_step(domain, block, dynamicsTimeStep, tracersHead)
!>
!>  The volume <-> thickness conversions around the block loop (:2462-2480, :2680-2700) are done by the
!>  library (volumeLike = 1), so the two loops that call volume_to_thickness / thickness_to_volume are
!>  skipped when the device path is on.  Halo updates before and after (:2410-2450, :2705-2712) stay as
!>  they are.  The optional checks: the library computes this block's conservation sums and
!>  check_tracer_conservation (:8126) keeps adding the ranks and testing them; the monotonicity test
!>  (:8416) runs in the library, so the call of check_tracer_monotonicity is skipped on the device path.
!>
!>  Not compiled in this repository (no Fortran compiler in the build image); tests/test_fortran_shim.py
!>  checks the interface blocks against include/ir_b200.h, and tests/test_fortran_shim_executed.py EXECUTES
!>  this module with an interpreter that calls the library through these interface blocks.
!
!-----------------------------------------------------------------------

module seaice_ir_b200

  use, intrinsic :: iso_c_binding

  use mpass_derived_types
  use mpass_pool_routines
  use mpass_log, only: mpas_log_write

  use seaice_advection_incremental_remap_tracers, only: &
       tracer_type, &
       seaice_set_tracer_array_pointers

  implicit none
  private
  save

  public :: &
       seaice_ir_b200_create, &
       seaice_ir_b200_step, &
       seaice_upwind_b200_step, &
       seaice_ir_b200_destroy

  ! ---- C ABI (include/ir_b200.h) ----------------------------------------------------------------

  type, bind(C) :: ir_mesh_desc
     integer(c_int) :: nCells, nCellsSolve, nVertices, nEdges, maxEdges, vertexDegree
     integer(c_int) :: nCategories
     integer(c_int) :: nQuadPoints
     integer(c_int) :: on_a_sphere
     integer(c_int) :: rotate_cartesian_grid
     type(c_ptr) :: nEdgesOnCell, edgesOnCell, cellsOnCell, verticesOnCell, cellsOnEdge, verticesOnEdge
     type(c_ptr) :: areaCell, dcEdge, coeffs_reconstruct
     type(c_ptr) :: transGlobalToCell
     type(c_ptr) :: xVertexOnCell, yVertexOnCell
     type(c_ptr) :: xVertexOnEdge, yVertexOnEdge
     type(c_ptr) :: remapEdge, cellsOnEdgeRemap, edgesOnEdgeRemap
     type(c_ptr) :: geomAvgCell(14)
  end type ir_mesh_desc

  type, bind(C) :: ir_tracer_desc
     integer(c_int) :: nLayers
     integer(c_int) :: parent
     integer(c_int) :: volumeLike
     type(c_ptr) :: array
  end type ir_tracer_desc

  type, bind(C) :: ir_upwind_var
     integer(c_int) :: parent
     integer(c_int) :: volumeLike
     real(c_double) :: childMinimum
     type(c_ptr) :: array
  end type ir_upwind_var

  type, bind(C) :: ir_check_report
     integer(c_int) :: conservationViolated
     integer(c_int) :: consTracer, consCategory, consLayer
     real(c_double) :: sumInit, sumFinal
     integer(c_int) :: monotonicityViolated
     integer(c_int) :: monoTracer, monoCategory, monoLayer, monoCell
     real(c_double) :: newValue, bound, tolerance
  end type ir_check_report

  interface

     function ir_create(handle, mesh, device) bind(C, name="ir_create") result(ierr)
       import :: c_ptr, c_int, ir_mesh_desc
       type(c_ptr), intent(out) :: handle
       type(ir_mesh_desc), intent(in) :: mesh
       integer(c_int), value :: device
       integer(c_int) :: ierr
     end function ir_create

     function ir_set_tracers(handle, nTracers, tracers) bind(C, name="ir_set_tracers") result(ierr)
       import :: c_ptr, c_int, ir_tracer_desc
       type(c_ptr), value :: handle
       integer(c_int), value :: nTracers
       type(ir_tracer_desc), dimension(*), intent(in) :: tracers
       integer(c_int) :: ierr
     end function ir_set_tracers

     function ir_run(handle, nTracers, tracers, uVelocity, vVelocity, dt) bind(C, name="ir_run") result(ierr)
       import :: c_ptr, c_int, c_double, ir_tracer_desc
       type(c_ptr), value :: handle
       integer(c_int), value :: nTracers
       type(ir_tracer_desc), dimension(*), intent(in) :: tracers
       type(c_ptr), value :: uVelocity, vVelocity
       real(c_double), value :: dt
       integer(c_int) :: ierr
     end function ir_run

     function ir_set_checks(handle, conservation, monotonicity) bind(C, name="ir_set_checks") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int), value :: conservation, monotonicity
       integer(c_int) :: ierr
     end function ir_set_checks

     function ir_fetch_check_report(handle, report) bind(C, name="ir_fetch_check_report") result(ierr)
       import :: c_ptr, c_int, ir_check_report
       type(c_ptr), value :: handle
       type(ir_check_report), intent(out) :: report
       integer(c_int) :: ierr
     end function ir_fetch_check_report

     function ir_fetch_conservation_sums(handle, tracer, sumInit, sumFinal) &
          bind(C, name="ir_fetch_conservation_sums") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int), value :: tracer
       type(c_ptr), value :: sumInit, sumFinal
       integer(c_int) :: ierr
     end function ir_fetch_conservation_sums

     function ir_set_upwind_mesh(handle, interiorEdge, dvEdge, normalVectorEdge) &
          bind(C, name="ir_set_upwind_mesh") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       type(c_ptr), value :: interiorEdge, dvEdge, normalVectorEdge
       integer(c_int) :: ierr
     end function ir_set_upwind_mesh

     function ir_run_upwind(handle, nVars, vars, uVelocity, vVelocity, dt) bind(C, name="ir_run_upwind") result(ierr)
       import :: c_ptr, c_int, c_double, ir_upwind_var
       type(c_ptr), value :: handle
       integer(c_int), value :: nVars
       type(ir_upwind_var), dimension(*), intent(in) :: vars
       type(c_ptr), value :: uVelocity, vVelocity
       real(c_double), value :: dt
       integer(c_int) :: ierr
     end function ir_run_upwind

     function ir_destroy(handle) bind(C, name="ir_destroy") result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: handle
       integer(c_int) :: ierr
     end function ir_destroy

     function ir_last_error_string() bind(C, name="ir_last_error_string") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function ir_last_error_string

  end interface

  integer(c_int), parameter :: IR_OK = 0
  integer(c_int), parameter :: IR_ERR_MONOTONICITY = 15

  ! one handle per block; MPAS-Seaice runs one block per rank on the device path (mesh_pool.F:98-105)
  type(c_ptr) :: irHandle = c_null_ptr

  integer, parameter :: maxTracers = 128
  type(ir_tracer_desc), dimension(maxTracers), target :: tracerTable
  integer :: nTracersTable = 0

  type(ir_upwind_var), dimension(4), target :: upwindTable
  logical :: upwindMeshSet = .false.

contains

!-----------------------------------------------------------------------
!  seaice_ir_b200_create: once, after seaice_init_advection_incremental_remap has filled the
!  incremental_remap pool (mpas_seaice_advection_incremental_remap.F:165-816)
!-----------------------------------------------------------------------

  subroutine seaice_ir_b200_create(block)

    type(block_type), intent(inout) :: block

    type(mpas_pool_type), pointer :: meshPool, incrementalRemapPool

    integer, pointer :: nCells, nCellsSolve, nVertices, nEdges, maxEdges, vertexDegree, nCategories, nQuadPoints
    logical, pointer :: on_a_sphere, config_rotate_cartesian_grid

    integer, dimension(:), pointer :: nEdgesOnCell, remapEdge
    integer, dimension(:,:), pointer :: edgesOnCell, cellsOnCell, verticesOnCell, cellsOnEdge, verticesOnEdge, &
         cellsOnEdgeRemap, edgesOnEdgeRemap
    real(kind=RKIND), dimension(:), pointer :: areaCell, dcEdge, geomAvg
    real(kind=RKIND), dimension(:,:), pointer :: xVertexOnCell, yVertexOnCell, xVertexOnEdge, yVertexOnEdge
    real(kind=RKIND), dimension(:,:,:), pointer :: coeffsReconstruct, transGlobalToCell

    character(len=12), dimension(14), parameter :: geomNames = (/ &
         'xAvgCell    ', 'yAvgCell    ', 'xxAvgCell   ', 'xyAvgCell   ', 'yyAvgCell   ', 'xxxAvgCell  ', 'xxyAvgCell  ', &
         'xyyAvgCell  ', 'yyyAvgCell  ', 'xxxxAvgCell ', 'xxxyAvgCell ', 'xxyyAvgCell ', 'xyyyAvgCell ', 'yyyyAvgCell ' /)

    type(ir_mesh_desc) :: mesh
    integer :: k
    integer(c_int) :: ierr

    call MPAS_pool_get_subpool(block % structs, 'mesh', meshPool)
    call MPAS_pool_get_subpool(block % structs, 'incremental_remap', incrementalRemapPool)

    call MPAS_pool_get_dimension(meshPool, 'nCells', nCells)
    call MPAS_pool_get_dimension(meshPool, 'nCellsSolve', nCellsSolve)
    call MPAS_pool_get_dimension(meshPool, 'nVertices', nVertices)
    call MPAS_pool_get_dimension(meshPool, 'nEdges', nEdges)
    call MPAS_pool_get_dimension(meshPool, 'maxEdges', maxEdges)
    call MPAS_pool_get_dimension(meshPool, 'vertexDegree', vertexDegree)
    call MPAS_pool_get_dimension(meshPool, 'nCategories', nCategories)
    call MPAS_pool_get_dimension(meshPool, 'nQuadPoints', nQuadPoints)
    call MPAS_pool_get_config(meshPool, 'on_a_sphere', on_a_sphere)
    call MPAS_pool_get_config(block % configs, 'config_rotate_cartesian_grid', config_rotate_cartesian_grid)

    call MPAS_pool_get_array(meshPool, 'nEdgesOnCell', nEdgesOnCell)
    call MPAS_pool_get_array(meshPool, 'edgesOnCell', edgesOnCell)
    call MPAS_pool_get_array(meshPool, 'cellsOnCell', cellsOnCell)
    call MPAS_pool_get_array(meshPool, 'verticesOnCell', verticesOnCell)
    call MPAS_pool_get_array(meshPool, 'cellsOnEdge', cellsOnEdge)
    call MPAS_pool_get_array(meshPool, 'verticesOnEdge', verticesOnEdge)
    call MPAS_pool_get_array(meshPool, 'areaCell', areaCell)
    call MPAS_pool_get_array(meshPool, 'dcEdge', dcEdge)
    call MPAS_pool_get_array(meshPool, 'coeffs_reconstruct', coeffsReconstruct)

    call MPAS_pool_get_array(incrementalRemapPool, 'transGlobalToCell', transGlobalToCell)
    call MPAS_pool_get_array(incrementalRemapPool, 'xVertexOnCell', xVertexOnCell)
    call MPAS_pool_get_array(incrementalRemapPool, 'yVertexOnCell', yVertexOnCell)
    call MPAS_pool_get_array(incrementalRemapPool, 'xVertexOnEdge', xVertexOnEdge)
    call MPAS_pool_get_array(incrementalRemapPool, 'yVertexOnEdge', yVertexOnEdge)
    call MPAS_pool_get_array(incrementalRemapPool, 'remapEdge', remapEdge)
    call MPAS_pool_get_array(incrementalRemapPool, 'cellsOnEdgeRemap', cellsOnEdgeRemap)
    call MPAS_pool_get_array(incrementalRemapPool, 'edgesOnEdgeRemap', edgesOnEdgeRemap)

    mesh % nCells = nCells
    mesh % nCellsSolve = nCellsSolve
    mesh % nVertices = nVertices
    mesh % nEdges = nEdges
    mesh % maxEdges = maxEdges
    mesh % vertexDegree = vertexDegree
    mesh % nCategories = nCategories
    mesh % nQuadPoints = nQuadPoints
    mesh % on_a_sphere = merge(1, 0, on_a_sphere)
    mesh % rotate_cartesian_grid = merge(1, 0, config_rotate_cartesian_grid)

    mesh % nEdgesOnCell = c_loc(nEdgesOnCell)
    mesh % edgesOnCell = c_loc(edgesOnCell)
    mesh % cellsOnCell = c_loc(cellsOnCell)
    mesh % verticesOnCell = c_loc(verticesOnCell)
    mesh % cellsOnEdge = c_loc(cellsOnEdge)
    mesh % verticesOnEdge = c_loc(verticesOnEdge)
    mesh % areaCell = c_loc(areaCell)
    mesh % dcEdge = c_loc(dcEdge)
    mesh % coeffs_reconstruct = c_loc(coeffsReconstruct)
    mesh % transGlobalToCell = c_loc(transGlobalToCell)
    mesh % xVertexOnCell = c_loc(xVertexOnCell)
    mesh % yVertexOnCell = c_loc(yVertexOnCell)
    mesh % xVertexOnEdge = c_loc(xVertexOnEdge)
    mesh % yVertexOnEdge = c_loc(yVertexOnEdge)
    mesh % remapEdge = c_loc(remapEdge)
    mesh % cellsOnEdgeRemap = c_loc(cellsOnEdgeRemap)
    mesh % edgesOnEdgeRemap = c_loc(edgesOnEdgeRemap)
    do k = 1, 14
       call MPAS_pool_get_array(incrementalRemapPool, trim(geomNames(k)), geomAvg)
       mesh % geomAvgCell(k) = c_loc(geomAvg)
    enddo

    ierr = ir_create(irHandle, mesh, -1_c_int)
    if (ierr /= IR_OK) call ir_b200_abort('ir_create')

  end subroutine seaice_ir_b200_create

!-----------------------------------------------------------------------
!  seaice_ir_b200_step: replaces incremental_remap_block (:2740) for one block
!-----------------------------------------------------------------------

  subroutine seaice_ir_b200_step(domain, block, dt, tracersHead)

    type(domain_type), intent(in) :: domain
    type(block_type), intent(inout) :: block
    real(kind=RKIND), intent(in) :: dt
    type(tracer_type), pointer :: tracersHead

    type(mpas_pool_type), pointer :: velocityPool
    real(kind=RKIND), dimension(:), pointer :: uVelocity, vVelocity

    type(tracer_type), pointer :: thisTracer, otherTracer
    integer :: nTracers, iTracer, iParent
    integer(c_int) :: ierr
    logical, pointer :: configConservationCheck, configMonotonicityCheck
    integer(c_int) :: consMode, monoMode

    call MPAS_pool_get_subpool(block % structs, 'velocity_solver', velocityPool)
    call MPAS_pool_get_array(velocityPool, 'uVelocity', uVelocity)
    call MPAS_pool_get_array(velocityPool, 'vVelocity', vVelocity)

    ! the ordered list: the mass-like field first, then the tracers level by level (set_tracer_order,
    ! incremental_remap_tracers.F:860-922), so a parent always precedes its children
    call seaice_set_tracer_array_pointers(tracersHead, block, 1)

    nTracers = 0
    thisTracer => tracersHead
    do while (associated(thisTracer))
       nTracers = nTracers + 1
       if (nTracers > maxTracers) call ir_b200_abort('more tracers than maxTracers')

       if (thisTracer % ndims == 2) then
          tracerTable(nTracers) % nLayers = 1
          tracerTable(nTracers) % array = c_loc(thisTracer % array2D)
       else
          tracerTable(nTracers) % nLayers = size(thisTracer % array3D, 1)
          tracerTable(nTracers) % array = c_loc(thisTracer % array3D)
       endif

       ! index (0-based) of the parent in the list
       tracerTable(nTracers) % parent = -1
       if (thisTracer % nParents > 0) then
          iParent = 0
          otherTracer => tracersHead
          do while (associated(otherTracer))
             if (associated(otherTracer, thisTracer % parent)) then
                tracerTable(nTracers) % parent = iParent
                exit
             endif
             iParent = iParent + 1
             otherTracer => otherTracer % next
          enddo
       endif

       ! volume on entry and exit, thickness while transported (:2462-2480, :2680-2700)
       tracerTable(nTracers) % volumeLike = 0
       if (trim(thisTracer % tracerName) == 'iceVolumeCategory' .or. &
           trim(thisTracer % tracerName) == 'snowVolumeCategory') tracerTable(nTracers) % volumeLike = 1

       thisTracer => thisTracer % next
    enddo

    ! (re)declare the hierarchy when the set of active tracers changed
    if (nTracers /= nTracersTable) then
       ierr = ir_set_tracers(irHandle, int(nTracers, c_int), tracerTable)
       if (ierr /= IR_OK) call ir_b200_abort('ir_set_tracers')
       nTracersTable = nTracers
    endif

    ! config_conservation_check / config_monotonicity_check (:2574, :2590).  Conservation in mode 2: the
    ! library computes this block's sums, check_tracer_conservation (:8126) keeps adding the ranks and testing.
    ! Monotonicity is tested by the library (every bound it needs lies within the halo layers check_halo_layer_number asks for).
    call MPAS_pool_get_config(block % configs, 'config_conservation_check', configConservationCheck)
    call MPAS_pool_get_config(block % configs, 'config_monotonicity_check', configMonotonicityCheck)
    consMode = 0
    monoMode = 0
    if (configConservationCheck) consMode = 2
    if (configMonotonicityCheck) monoMode = 1
    ierr = ir_set_checks(irHandle, consMode, monoMode)
    if (ierr /= IR_OK) call ir_b200_abort('ir_set_checks')

    ierr = ir_run(irHandle, int(nTracers, c_int), tracerTable, c_loc(uVelocity), c_loc(vVelocity), real(dt, c_double))
    if (ierr /= IR_OK) call ir_b200_abort('ir_run')

    ! the sums go where check_tracer_conservation reads them (incremental_remap_tracers.F:54-57)
    if (configConservationCheck) then
       iTracer = 0
       thisTracer => tracersHead
       do while (associated(thisTracer))
          if (thisTracer % ndims == 2) then
             ierr = ir_fetch_conservation_sums(irHandle, int(iTracer, c_int), &
                  c_loc(thisTracer % globalSumInit2D), c_loc(thisTracer % globalSumFinal2D))
          else
             ierr = ir_fetch_conservation_sums(irHandle, int(iTracer, c_int), &
                  c_loc(thisTracer % globalSumInit3D), c_loc(thisTracer % globalSumFinal3D))
          endif
          if (ierr /= IR_OK) call ir_b200_abort('ir_fetch_conservation_sums')
          iTracer = iTracer + 1
          thisTracer => thisTracer % next
       enddo
    endif

  end subroutine seaice_ir_b200_step

!-----------------------------------------------------------------------
!  seaice_upwind_b200_step: config_advection_type = 'upwind'.  Replaces, for one block, what
!  seaice_run_advection_upwind (mpas_seaice_advection_upwind.F:385-520) does between the tracer halo
!  exchange and the time-level shift: prepare_advection, the loop over tracerConnectivities,
!  finalize_advection.  The table is the reference's (define_tracer_connectivities :145-170), in its
!  order; the library works in place on time level 1, so MPAS_pool_shift_time_levels (:2032) is
!  skipped for these four arrays.  (The reference zeroes time level 2 of iceEnthalpy, iceSalinity and
!  snowEnthalpy (:1690-1700) and shifts it in: a host that wants that too zeroes them here.)
!-----------------------------------------------------------------------

  subroutine seaice_upwind_b200_step(block, dt)

    type(block_type), intent(inout) :: block
    real(kind=RKIND), intent(in) :: dt

    type(mpas_pool_type), pointer :: meshPool, boundaryPool, tracersPool, velocityPool
    real(kind=RKIND), dimension(:), pointer :: uVelocity, vVelocity, dvEdge
    real(kind=RKIND), dimension(:,:,:), pointer :: normalVectorEdge
    real(kind=RKIND), dimension(:,:,:), pointer :: iceAreaCategory, surfaceTemperature, iceVolumeCategory, snowVolumeCategory
    integer, dimension(:), pointer :: interiorEdge
    integer(c_int) :: ierr

    call MPAS_pool_get_subpool(block % structs, 'mesh', meshPool)
    call MPAS_pool_get_subpool(block % structs, 'boundary', boundaryPool)
    call MPAS_pool_get_subpool(block % structs, 'tracers', tracersPool)
    call MPAS_pool_get_subpool(block % structs, 'velocity_solver', velocityPool)

    call MPAS_pool_get_array(velocityPool, 'uVelocity', uVelocity)
    call MPAS_pool_get_array(velocityPool, 'vVelocity', vVelocity)

    if (.not. upwindMeshSet) then
       call MPAS_pool_get_array(meshPool, 'dvEdge', dvEdge)
       call MPAS_pool_get_array(boundaryPool, 'interiorEdge', interiorEdge)
       call MPAS_pool_get_array(velocityPool, 'normalVectorEdge', normalVectorEdge)
       ierr = ir_set_upwind_mesh(irHandle, c_loc(interiorEdge), c_loc(dvEdge), c_loc(normalVectorEdge))
       if (ierr /= IR_OK) call ir_b200_abort('ir_set_upwind_mesh')
       upwindMeshSet = .true.
    endif

    call MPAS_pool_get_array(tracersPool, 'iceAreaCategory', iceAreaCategory, 1)
    call MPAS_pool_get_array(tracersPool, 'surfaceTemperature', surfaceTemperature, 1)
    call MPAS_pool_get_array(tracersPool, 'iceVolumeCategory', iceVolumeCategory, 1)
    call MPAS_pool_get_array(tracersPool, 'snowVolumeCategory', snowVolumeCategory, 1)

    ! (1, nCategories, nCells+1) arrays: the same memory as the (nCategories, nCells+1) the C side takes
    upwindTable(1) % parent = -1
    upwindTable(1) % volumeLike = 0
    upwindTable(1) % array = c_loc(iceAreaCategory)
    upwindTable(2) % parent = 0
    upwindTable(2) % volumeLike = 0
    upwindTable(2) % array = c_loc(surfaceTemperature)
    upwindTable(3) % parent = 1
    upwindTable(3) % volumeLike = 1
    upwindTable(3) % array = c_loc(iceVolumeCategory)
    upwindTable(4) % parent = 2
    upwindTable(4) % volumeLike = 1
    upwindTable(4) % array = c_loc(snowVolumeCategory)
    upwindTable(:) % childMinimum = 0.0_c_double

    ierr = ir_run_upwind(irHandle, 4_c_int, upwindTable, c_loc(uVelocity), c_loc(vVelocity), real(dt, c_double))
    if (ierr /= IR_OK) call ir_b200_abort('ir_run_upwind')

  end subroutine seaice_upwind_b200_step

!-----------------------------------------------------------------------

  subroutine seaice_ir_b200_destroy()

    integer(c_int) :: ierr

    if (c_associated(irHandle)) then
       ierr = ir_destroy(irHandle)
       irHandle = c_null_ptr
       nTracersTable = 0
       upwindMeshSet = .false.
    endif

  end subroutine seaice_ir_b200_destroy

!-----------------------------------------------------------------------

  subroutine ir_b200_abort(where)

    character(len=*), intent(in) :: where

    character(kind=c_char), dimension(:), pointer :: cmsg
    character(len=512) :: msg
    type(c_ptr) :: p
    integer :: i

    msg = ''
    p = ir_last_error_string()
    if (c_associated(p)) then
       call c_f_pointer(p, cmsg, (/512/))
       do i = 1, 512
          if (cmsg(i) == c_null_char) exit
          msg(i:i) = cmsg(i)
       enddo
    endif
    call mpas_log_write('seaice_ir_b200: '//trim(where)//' failed: '//trim(msg), MPAS_LOG_CRIT)

  end subroutine ir_b200_abort

end module seaice_ir_b200
