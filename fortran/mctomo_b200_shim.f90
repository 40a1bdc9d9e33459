! mctomo_b200_shim.f90 -- ISO_C_BINDING glue between the unchanged MCTomo host and libmctomo_b200.so.
!
! This file is SOURCE ONLY: the build image of this repository has no Fortran compiler, so it has not
! been compiled here (needs the preprocessor: name it .F90 or pass -cpp, as the reference does for its .F90 files).  It is written against the reference's own types and mirrors the style of the
! reference's existing C wrappers (src/fastMarching_wrapper.f90:18-74, src/cgal_delaunay_wrapper.f90:51-211).
!
! What a maintainer does (INTEGRATION.md has the step-by-step):
!   1. add this file to the object list of src/makefile and link with -lmctomo_b200 -lcudart;
!   2. call mctomo_b200_init(rank) once after mpi_init (src/MCTomo.F90:82-86);
!   3. replace the BODY of kdtree_to_grid (src/mcmc_loc2.f90:2029-2078) by
!          call kdtree_to_grid_b200(RTI, grid, bnd_box, model, pm)
!      -- the subroutine statement, argument list and callers stay as they are;
!   4. replace lines 161-206 of surf_likelihood (src/likelihood_surf.F90) by
!          call surf_dispersion_b200(model, grid, ix0, ix1, iy0, iy1, dat%freqs, settings%raylov, &
!                                    settings%phaseGroup, 0, settings%dPhaseVel, pvel, gvel, ierr, invalid)
!      and `if (invalid) then; like%like = huge(like%like); return; endif` (the index window code of
!      lines 155-158,166-169 stays in front of it);
!   5. same for program modelling (src/forward_modelling.f90:393-429) with variant = 1.
module m_mctomo_b200
    use iso_c_binding
    use m_settings, only : T_GRID, T_MOD
    use run_info,   only : T_RUN_INFO
    use cgal_delaunay, only : d3, p3
    implicit none
    private

    public :: mctomo_b200_init, mctomo_b200_shutdown
    public :: kdtree_to_grid_b200, surf_dispersion_b200, vs2vp_rho_b200
    ! the fast-marching travel times of every (period, source) in one call (INTEGRATION.md 6)
    public :: fm2d_times_b200, fm2d_rays_b200
    ! the resident session: one chain's model kept in HBM between proposals (INTEGRATION.md 5a)
    public :: T_B200_SESSION, b200_session_create, b200_session_destroy, b200_session_set_model, b200_session_propose, &
              b200_session_accept, b200_session_reject, b200_session_likelihood, b200_session_stat_rti
    ! one chain on several GPUs (BASELINE config 5): NCCL communicator bootstrap over MPI
    public :: mctomo_b200_comm_init

    ! handle + host staging of one session
    type T_B200_SESSION
        type(c_ptr) :: h = c_null_ptr
        integer :: nx = 0, ny = 0, nout = 0
        real(c_double), allocatable :: pvel_w(:), gvel_w(:)      ! packed window maps of the last proposal
        integer(c_int), allocatable :: ierr_w(:)
        integer(c_int) :: win(4) = 0                             ! ix0, ix1, iy0, iy1 of the last proposal (box + halo)
    end type

    ! mirrors `mct_grid` of include/mctomo_b200.h
    type, bind(C) :: mct_grid
        integer(c_int32_t) :: nx, ny, nz
        real(c_double)     :: xmin, ymin, zmin
        real(c_double)     :: dx, dy, dz
        real(c_double)     :: waterDepth
        real(c_double)     :: scaling
    end type

    ! mirrors `mct_fm2d_opts`
    type, bind(C) :: mct_fm2d_opts
        integer(c_int32_t) :: gridx, gridy, sgref, sgdic, sgext, order
        real(c_double)     :: band
    end type

    ! mirrors `mct_disp_opts`
    type, bind(C) :: mct_disp_opts
        integer(c_int32_t) :: raylov, phaseGroup, nmodes, check_scope
        real(c_double)     :: dphase, layer_eps, water_thresh, preset
    end type

    interface
        integer(c_int) function mct_init(device) bind(C, name='mct_init')
            import :: c_int
            integer(c_int), value :: device
        end function
        integer(c_int) function mct_shutdown() bind(C, name='mct_shutdown')
            import :: c_int
        end function
        function mct_last_error() bind(C, name='mct_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        integer(c_int) function mct_voronoi_to_grid(points, params, ncells, g, box, pm, vp, vs, rho, sites_id) &
                bind(C, name='mct_voronoi_to_grid')
            import :: c_int, c_ptr, mct_grid
            type(c_ptr), value    :: points, params      ! real(c_double) (3,ncells)
            integer(c_int), value :: ncells
            type(mct_grid), intent(in) :: g
            type(c_ptr), value    :: box                 ! real(c_double) (6): x0,y0,z0,x1,y1,z1
            type(c_ptr), value    :: pm                  ! c_null_ptr or real(c_double) (3): vp,vs,rho
            type(c_ptr), value    :: vp, vs, rho         ! real(c_double) (nz,ny,nx)
            type(c_ptr), value    :: sites_id            ! integer(c_int) (nz,ny,nx)
        end function
        integer(c_int) function mct_vs2vp_rho(vs, vp, rho, n) bind(C, name='mct_vs2vp_rho')
            import :: c_int, c_ptr, c_int64_t
            type(c_ptr), value :: vs, vp, rho
            integer(c_int64_t), value :: n
        end function
        integer(c_int) function mct_vs2vp_rho_window(vs, vp, rho, g, w) bind(C, name='mct_vs2vp_rho_window')
            import :: c_int, c_ptr, mct_grid
            type(c_ptr), value :: vs, vp, rho
            type(mct_grid), intent(in) :: g
            type(c_ptr), value :: w                      ! integer(c_int32_t) (6): ix0,ix1,iy0,iy1,iz0,iz1
        end function
        integer(c_int) function mct_box_window(g, box, w) bind(C, name='mct_box_window')
            import :: c_int, c_ptr, mct_grid
            type(mct_grid), intent(in) :: g
            type(c_ptr), value :: box, w
        end function
        integer(c_int) function mct_surf_dispersion(vp, vs, rho, g, ix0, ix1, iy0, iy1, freqs, np, opt, &
                pvel, gvel, ierr, model_invalid) bind(C, name='mct_surf_dispersion')
            import :: c_int, c_ptr, mct_grid, mct_disp_opts
            type(c_ptr), value    :: vp, vs, rho
            type(mct_grid), intent(in) :: g
            integer(c_int), value :: ix0, ix1, iy0, iy1
            type(c_ptr), value    :: freqs
            integer(c_int), value :: np
            type(mct_disp_opts), intent(in) :: opt
            type(c_ptr), value    :: pvel, gvel, ierr
            type(c_ptr), value    :: model_invalid       ! c_null_ptr = skip check_model
        end function
        ! ---- resident session (INTEGRATION.md 5a): the chain's model stays in HBM between proposals ----
        integer(c_int) function mct_session_create(g, freqs, np, opt, derive_vp_rho, sess) bind(C, name='mct_session_create')
            import :: c_int, c_ptr, mct_grid, mct_disp_opts
            type(mct_grid), intent(in) :: g
            type(c_ptr), value    :: freqs
            integer(c_int), value :: np
            type(mct_disp_opts), intent(in) :: opt
            integer(c_int), value :: derive_vp_rho
            type(c_ptr), intent(out) :: sess
        end function
        integer(c_int) function mct_session_destroy(sess) bind(C, name='mct_session_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess
        end function
        integer(c_int) function mct_session_set_model(sess, points, params, ncells, pvel, gvel, ierr, model_invalid) &
                bind(C, name='mct_session_set_model')
            import :: c_int, c_ptr
            type(c_ptr), value    :: sess, points, params
            integer(c_int), value :: ncells
            type(c_ptr), value    :: pvel, gvel, ierr, model_invalid   ! c_null_ptr: keep on the device
        end function
        integer(c_int) function mct_session_propose(sess, points, params, ncells, box, pm, win, pvel_w, gvel_w, ierr_w, &
                model_invalid) bind(C, name='mct_session_propose')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value    :: sess, points, params
            integer(c_int), value :: ncells
            real(c_double), intent(in) :: box(6)
            type(c_ptr), value    :: pm                              ! c_null_ptr, or (vp,vs,rho) of a value-only move
            integer(c_int), intent(out) :: win(4)                    ! ix0, ix1, iy0, iy1 (1-based, box + halo)
            type(c_ptr), value    :: pvel_w, gvel_w, ierr_w          ! packed (nout,wy,wx) / (wy,wx); room for nx*ny columns
            integer(c_int), intent(out) :: model_invalid
        end function
        integer(c_int) function mct_session_accept(sess) bind(C, name='mct_session_accept')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess
        end function
        integer(c_int) function mct_session_reject(sess) bind(C, name='mct_session_reject')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess
        end function
        integer(c_int) function mct_session_group_times(sess, ray_points, ray_offsets, nrays, time) &
                bind(C, name='mct_session_group_times')
            import :: c_int, c_ptr
            type(c_ptr), value    :: sess
            type(c_ptr), value    :: ray_points      ! real(c_double) (2, total points), rays packed back to back
            type(c_ptr), value    :: ray_offsets     ! integer(c_int64_t) (nrays*np + 1), 0-based first point of each ray
            integer(c_int), value :: nrays
            type(c_ptr), value    :: time            ! real(c_double) (nrays, np) out = like%phaseTime flattened over (nrev,nsrc)
        end function
        integer(c_int) function mct_session_get_model(sess, vp, vs, rho, sites_id) bind(C, name='mct_session_get_model')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess, vp, vs, rho, sites_id
        end function
        integer(c_int) function mct_session_get_maps(sess, pvel, gvel, ierr) bind(C, name='mct_session_get_maps')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess, pvel, gvel, ierr
        end function
        integer(c_int) function mct_session_set_rays(sess, ray_points, ray_offsets, nrays) bind(C, name='mct_session_set_rays')
            import :: c_int, c_ptr
            type(c_ptr), value    :: sess, ray_points, ray_offsets
            integer(c_int), value :: nrays
        end function
        integer(c_int) function mct_session_set_data(sess, nrr, sigdep, nrays_total, ttime, raystat, srdist) &
                bind(C, name='mct_session_set_data')
            import :: c_int, c_ptr
            type(c_ptr), value    :: sess
            integer(c_int), value :: nrr, sigdep, nrays_total
            type(c_ptr), value    :: ttime, raystat, srdist      ! dat%ttime (nrr,3,np), dat%raystat (nrr,2,np), like%srdist (nrr,np)
        end function
        integer(c_int) function mct_session_likelihood(sess, pending, ray_points, ray_offsets, nrays, snoise0, snoise1, &
                out, phase_time, sigma) bind(C, name='mct_session_likelihood')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value    :: sess
            integer(c_int), value :: pending                     ! 0: current model, 1: pending proposal
            type(c_ptr), value    :: ray_points, ray_offsets     ! c_null_ptr: the rays of mct_session_set_rays
            integer(c_int), value :: nrays
            type(c_ptr), value    :: snoise0, snoise1            ! RTI%snoise0, RTI%snoise1 (np) or c_null_ptr (sigdep == 0)
            real(c_double), intent(out) :: out(3)                ! like%like, like%misfit, like%unweighted_misfit
            type(c_ptr), value    :: phase_time, sigma           ! optional (nrr,np) outputs, c_null_ptr to skip
        end function
        integer(c_int) function mct_surf_misfit(time, nrr, np, sigdep, nrays_total, ttime, raystat, snoise0, snoise1, srdist, &
                out, sigma) bind(C, name='mct_surf_misfit')
            import :: c_int, c_ptr, c_double
            type(c_ptr), value    :: time
            integer(c_int), value :: nrr, np, sigdep, nrays_total
            type(c_ptr), value    :: ttime, raystat, snoise0, snoise1, srdist
            real(c_double), intent(out) :: out(3)
            type(c_ptr), value    :: sigma
        end function
        integer(c_int) function mct_session_stat_accumulate(sess) bind(C, name='mct_session_stat_accumulate')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess
        end function
        integer(c_int) function mct_session_stat_get(sess, aveS, stdS, aveP, stdP, nsamples) bind(C, name='mct_session_stat_get')
            import :: c_int, c_ptr
            type(c_ptr), value :: sess, aveS, stdS, aveP, stdP, nsamples
        end function
        ! ---- 2-D fast-marching travel times (k6_fm2d.cuh): modrays for phase-velocity data, all periods x sources at once ----
        integer(c_int) function mct_fm2d_times(src_x, src_z, nsrc, rcv_x, rcv_z, nrc, srs, vel, nmaps, nvx, nvz, gox, goz, dvx, dvz, &
                opt, ttime, field) bind(C, name='mct_fm2d_times')
            import :: c_int, c_ptr, c_double, mct_fm2d_opts
            type(c_ptr), value    :: src_x, src_z, rcv_x, rcv_z, srs, vel, ttime, field
            integer(c_int), value :: nsrc, nrc, nmaps, nvx, nvz
            real(c_double), value :: gox, goz, dvx, dvz
            type(mct_fm2d_opts), intent(in) :: opt
        end function
        integer(c_int) function mct_fm2d_rays(src_x, src_z, nsrc, rcv_x, rcv_z, nrc, srs, srsv, vel, nmaps, nvx, nvz, gox, goz, dvx, dvz, &
                opt, ttime, ray_cap, ray_npts, ray_pts, ray_len, crazy) bind(C, name='mct_fm2d_rays')
            import :: c_int, c_ptr, c_double, mct_fm2d_opts
            type(c_ptr), value    :: src_x, src_z, rcv_x, rcv_z, srs, srsv, vel, ttime, ray_npts, ray_pts, ray_len, crazy
            integer(c_int), value :: nsrc, nrc, nmaps, nvx, nvz, ray_cap
            real(c_double), value :: gox, goz, dvx, dvz
            type(mct_fm2d_opts), intent(in) :: opt
        end function
        ! ---- low-velocity columns: the generalized R/T branch of surfmodes on the device (k5_grt.cuh) ----
        integer(c_int) function mct_set_grt(enable, par6) bind(C, name='mct_set_grt')
            import :: c_int, c_ptr
            integer(c_int), value :: enable
            type(c_ptr), value    :: par6        ! {tolmin, tolmax, smin_min, smin_max, dcm, dc2} of T_MODES_PARA, or NULL
        end function
        ! ---- one chain on several GPUs: NCCL data plane inside the library (include/mctomo_b200.h, "multi-GPU") ----
        integer(c_int) function mct_comm_unique_id(id128) bind(C, name='mct_comm_unique_id')
            import :: c_int, c_ptr
            type(c_ptr), value :: id128                          ! 128 bytes out
        end function
        integer(c_int) function mct_comm_init(id128, rank, nranks) bind(C, name='mct_comm_init')
            import :: c_int, c_ptr
            type(c_ptr), value    :: id128
            integer(c_int), value :: rank, nranks
        end function
        integer(c_int) function mct_comm_destroy() bind(C, name='mct_comm_destroy')
            import :: c_int
        end function
        integer(c_int) function mct_forward_sharded_dev(g, derive_vp_rho, freqs, np, opt, d_vp, d_vs, d_rho, d_sites, &
                d_pvel, d_gvel, d_ierr, d_flags, stream) bind(C, name='mct_forward_sharded_dev')
            import :: c_int, c_ptr, mct_grid, mct_disp_opts
            type(mct_grid), intent(in) :: g
            integer(c_int), value :: derive_vp_rho
            type(c_ptr), value    :: freqs
            integer(c_int), value :: np
            type(mct_disp_opts), intent(in) :: opt
            type(c_ptr), value    :: d_vp, d_vs, d_rho, d_sites, d_pvel, d_gvel, d_ierr, d_flags, stream   ! DEVICE pointers
        end function
    end interface

contains

    ! chains are MPI ranks (src/MCTomo.F90:131-133); several ranks may share a GPU.  local_rank: the rank within the node
    ! (MPI_Comm_split_type(MPI_COMM_TYPE_SHARED) / OMPI_COMM_WORLD_LOCAL_RANK / SLURM_LOCALID); ngpus: devices of the node
    ! (cudaGetDeviceCount, or the launcher's setting).  Without them the global rank and 8 devices are assumed.
    subroutine mctomo_b200_init(rank, local_rank, ngpus)
        integer, intent(in) :: rank
        integer, intent(in), optional :: local_rank, ngpus
        integer(c_int) :: rc
        integer :: lr, ng
        lr = rank
        if (present(local_rank)) lr = local_rank
        ng = 8
        if (present(ngpus)) ng = max(1, ngpus)
        rc = mct_init(int(mod(lr, ng), c_int))
        if (rc /= 0) call fail('mct_init', rc)
    end subroutine

    ! NCCL communicator for the column-sharded evaluation of ONE chain: rank 0 draws the 128-byte id, MPI moves it
    ! (the reference already initialises MPI, src/MCTomo.F90:82-86).  Compile with -DMPI like the reference.
    subroutine mctomo_b200_comm_init(comm_rank, comm_size, mpi_comm)
        integer, intent(in) :: comm_rank, comm_size, mpi_comm
        integer(c_int8_t), target :: id(128)
        integer(c_int) :: rc
        integer :: ierror
        id = 0
        if (comm_rank == 0) then
            rc = mct_comm_unique_id(c_loc(id))
            if (rc /= 0) call fail('mct_comm_unique_id', rc)
        endif
#ifdef MPI
        call mpi_bcast(id, 128, MPI_BYTE, 0, mpi_comm, ierror)
#endif
        rc = mct_comm_init(c_loc(id), int(comm_rank, c_int), int(comm_size, c_int))
        if (rc /= 0) call fail('mct_comm_init', rc)
    end subroutine

    subroutine mctomo_b200_shutdown()
        integer(c_int) :: rc
        rc = mct_shutdown()
    end subroutine

    function c_grid(grid) result(g)
        type(T_GRID), intent(in) :: grid
        type(mct_grid) :: g
        g%nx = grid%nx; g%ny = grid%ny; g%nz = grid%nz
        g%xmin = grid%xmin; g%ymin = grid%ymin; g%zmin = grid%zmin
        g%dx = grid%dx; g%dy = grid%dy; g%dz = grid%dz
        g%waterDepth = grid%waterDepth
        g%scaling = grid%scaling
    end function

    ! Drop-in body for kdtree_to_grid (same dummy arguments as src/mcmc_loc2.f90:2002-2013).
    subroutine kdtree_to_grid_b200(RTI, grid, bnd_box, model, pm)
        type(T_RUN_INFO), intent(inout), target :: RTI
        type(T_GRID), intent(in) :: grid
        type(d3), dimension(2), intent(in) :: bnd_box
        type(T_MOD), intent(inout), target :: model
        type(p3), intent(in), optional :: pm

        real(c_double), target :: box(6), pmv(3)
        type(mct_grid) :: g
        type(c_ptr) :: pm_ptr
        integer(c_int) :: rc

        g = c_grid(grid)
        box = [bnd_box(1)%x, bnd_box(1)%y, bnd_box(1)%z, bnd_box(2)%x, bnd_box(2)%y, bnd_box(2)%z]
        pm_ptr = c_null_ptr
        if (present(pm)) then
            pmv = [pm%vp, pm%vs, pm%rho]
            pm_ptr = c_loc(pmv)
        endif
        ! points/parameters are (3,ncell_max): the first ncells columns are contiguous
        rc = mct_voronoi_to_grid(c_loc(RTI%points), c_loc(RTI%parameters), int(RTI%ncells, c_int), g, &
                                 c_loc(box), pm_ptr, c_loc(model%vp), c_loc(model%vs), c_loc(model%rho), &
                                 c_loc(RTI%sites_id))
        if (rc /= 0) call fail('mct_voronoi_to_grid', rc)
    end subroutine

    ! vs2vp_3d + vp2rho_3d (src/likelihood.f90:75-76).  With bnd_box present only the nodes kdtree_to_grid just
    ! rewrote are recomputed -- the rest of vp, rho is unchanged since the previous call, so the result equals the
    ! reference's whole-grid pass at a fraction of the transfers.
    subroutine vs2vp_rho_b200(model, grid, bnd_box)
        type(T_MOD), intent(inout), target :: model
        type(T_GRID), intent(in), optional :: grid
        type(d3), dimension(2), intent(in), optional :: bnd_box
        real(c_double), target :: box(6)
        integer(c_int32_t), target :: w(6)
        type(mct_grid) :: g
        integer(c_int) :: rc
        if (present(grid) .and. present(bnd_box)) then
            g = c_grid(grid)
            box = [bnd_box(1)%x, bnd_box(1)%y, bnd_box(1)%z, bnd_box(2)%x, bnd_box(2)%y, bnd_box(2)%z]
            rc = mct_box_window(g, c_loc(box), c_loc(w))
            if (rc == 0) rc = mct_vs2vp_rho_window(c_loc(model%vs), c_loc(model%vp), c_loc(model%rho), g, c_loc(w))
        else
            rc = mct_vs2vp_rho(c_loc(model%vs), c_loc(model%vp), c_loc(model%rho), int(size(model%vs), c_int64_t))
        endif
        if (rc /= 0) call fail('mct_vs2vp_rho', rc)
    end subroutine

    ! Replaces check_model + convert_to_layer + the OpenMP surfmodes loop (src/likelihood_surf.F90:161-206).
    ! variant 0: likelihood_surf.F90 constants (EPS = 1E-10, /scaling, presets 100);
    ! variant 1: forward_modelling.f90 constants (EPS = 1E-5, presets 1000, no check_model).
    ! pvel, gvel: (np*max(nmodes,1), iy0:iy1, ix0:ix1); ierr: (iy0:iy1, ix0:ix1)
    subroutine surf_dispersion_b200(model, grid, ix0, ix1, iy0, iy1, freqs, raylov, phaseGroup, nmodes, &
                                    dPhaseVel, pvel, gvel, ierr, invalid, variant, tol)
        type(T_MOD), intent(in), target :: model
        type(T_GRID), intent(in) :: grid
        integer, intent(in) :: ix0, ix1, iy0, iy1
        real(c_double), dimension(:), intent(in), target :: freqs
        integer, intent(in) :: raylov, phaseGroup, nmodes
        real(c_double), intent(in) :: dPhaseVel
        real(c_double), dimension(:,:,:), intent(inout), target :: pvel, gvel
        integer(c_int), dimension(:,:), intent(inout), target :: ierr
        logical, intent(out) :: invalid
        integer, intent(in), optional :: variant
        real(c_double), intent(in), optional :: tol           ! settings%tol: only the generalized R/T branch reads it

        type(mct_grid) :: g
        type(mct_disp_opts) :: opt
        integer(c_int), target :: inval
        integer(c_int) :: rc
        integer :: var
        real(c_double), target :: grtpar(6)

        var = 0
        if (present(variant)) var = variant
        g = c_grid(grid)
        opt%raylov = raylov
        opt%phaseGroup = phaseGroup
        opt%nmodes = nmodes
        ! 0 = check_model over the whole grid, exactly as the reference; 1 = over the window only (equivalent in the
        ! sampler, where the model outside the perturbed box is the previously accepted, valid one) and cheaper
        opt%check_scope = 0
        opt%dphase = dPhaseVel
        if (var == 0) then
            opt%layer_eps = real(1.0E-10, c_double)      ! EPS, likelihood_surf.F90:37 (a default-real literal)
            opt%water_thresh = real(1.0E-10, c_double)
            opt%preset = 100.0_c_double
        else
            opt%layer_eps = real(1.0E-5, c_double)       ! EPS, forward_modelling.f90:27
            opt%water_thresh = 0.0_c_double
            opt%preset = 1000.0_c_double
        endif
        ! T_MODES_PARA of the generalized R/T branch, as the two callers set it: likelihood_surf.F90:175-182
        ! (tolmin = settings%tol, tolmax = 10*tol) and forward_modelling.f90:397-404 (1E-6, 1e-7); smin and the steps are
        ! default-real literals there, hence real(.., c_double) of a single-precision constant.  Columns with a
        ! low-velocity layer are then solved on the device (surfmodes.f90:84-87,96-99), not reported as ierr = 2.
        if (var == 0) then
            grtpar(1) = 1.0E-6_c_double
            if (present(tol)) grtpar(1) = tol
            grtpar(2) = 10 * grtpar(1)
            grtpar(5) = dPhaseVel; grtpar(6) = dPhaseVel
        else
            grtpar(1) = real(1E-6, c_double); grtpar(2) = real(1e-7, c_double)
            grtpar(5) = real(1E-3, c_double); grtpar(6) = real(1E-3, c_double)
        endif
        grtpar(3) = real(1E-3, c_double); grtpar(4) = real(5E-3, c_double)
        rc = mct_set_grt(1_c_int, c_loc(grtpar))
        if (rc /= 0) call fail('mct_set_grt', rc)
        inval = 0
        if (var == 0) then
            rc = mct_surf_dispersion(c_loc(model%vp), c_loc(model%vs), c_loc(model%rho), g, ix0, ix1, iy0, iy1, &
                                     c_loc(freqs), size(freqs), opt, c_loc(pvel), c_loc(gvel), c_loc(ierr), c_loc(inval))
        else
            rc = mct_surf_dispersion(c_loc(model%vp), c_loc(model%vs), c_loc(model%rho), g, ix0, ix1, iy0, iy1, &
                                     c_loc(freqs), size(freqs), opt, c_loc(pvel), c_loc(gvel), c_loc(ierr), c_null_ptr)
        endif
        invalid = (inval /= 0)
        if (rc < 0 .or. rc == 1) call fail('mct_surf_dispersion', rc)
        ! rc = 2 (MCT_E_GRT_NEEDED) can only remain for a column the device branch does not cover (a multi-mode call:
        ! surfmmodes prints "not supported yet" there, surfmodes.f90:153,165; more than one fluid layer): such columns
        ! carry ierr = 2 and the preset velocities and are handed to the Fortran surfmodes here, column by column.
        ! rc = 3 / 5 (more than 200 layers / a fluid layer below the top): the reference overruns its arrays or `stop`s
        ! (surfmodes.f90:342-345); raise instead of carrying bogus velocities into fm2d.
        if (rc == 3 .or. rc == 5) call fail('mct_surf_dispersion', rc)
        if (rc == 2) call solve_lvl_columns_on_host(model, grid, ix0, ix1, iy0, iy1, freqs, raylov, phaseGroup, dPhaseVel, &
                                                    var, pvel, gvel, ierr)
    end subroutine

    ! Drop-in for the OpenMP loop over periods around `modrays` in surf_likelihood (src/likelihood_surf.F90:295-336) when
    ! settings%phaseGroup == 0 (uar = 1: travel times only).  src(2,nsrc), rev(2,nrev), raystat(nrev*nsrc,2,np),
    ! vel(np,ny+2,nx+2) = like%vel, phaseTime(nrev,nsrc,np) = like%phaseTime exactly as the reference declares them.
    ! Group-velocity data need the ray geometry (rpaths): fm2d_rays_b200 below.  Status 7 (MCT_E_FM2D_STALE: a source inside the
    ! model's last cell row/column, where the reference's refined march dies and returns the previous source's field) is raised
    ! like any other error: move the grid edge, or keep modrays for such a geometry.
    subroutine fm2d_times_b200(src, rev, raystat, grid, vel, gridx, gridy, sgref, sgdic, sgext, order, band, phaseTime)
        real(c_double), dimension(:,:), intent(in) :: src, rev
        integer(c_int), dimension(:,:,:), intent(in) :: raystat
        type(T_GRID), intent(in) :: grid
        real(c_double), dimension(:,:,:), intent(in) :: vel
        integer, intent(in) :: gridx, gridy, sgref, sgdic, sgext, order
        real(c_double), intent(in) :: band
        real(c_double), dimension(:,:,:), intent(inout), target :: phaseTime

        real(c_double), allocatable, target :: sx(:), sz(:), rx(:), rz(:), maps(:,:,:)
        integer(c_int), allocatable, target :: srs(:,:)
        type(mct_fm2d_opts) :: o
        integer(c_int) :: rc
        integer :: i, np, nsrc, nrev

        nsrc = size(src, 2); nrev = size(rev, 2); np = size(vel, 1)
        allocate(sx(nsrc), sz(nsrc), rx(nrev), rz(nrev), srs(nrev*nsrc, np), maps(size(vel, 2), size(vel, 3), np))
        sx = src(1,:); sz = src(2,:); rx = rev(1,:); rz = rev(2,:)
        srs = raystat(:, 1, :)
        do i = 1, np
            maps(:, :, i) = vel(i, :, :)
        enddo
        o%gridx = gridx; o%gridy = gridy; o%sgref = sgref; o%sgdic = sgdic; o%sgext = sgext; o%order = order; o%band = band
        rc = mct_fm2d_times(c_loc(sx), c_loc(sz), int(nsrc, c_int), c_loc(rx), c_loc(rz), int(nrev, c_int), c_loc(srs), c_loc(maps), &
                            int(np, c_int), int(grid%nx, c_int), int(grid%ny, c_int), grid%xmin, grid%ymin, grid%dx, grid%dy, o, &
                            c_loc(phaseTime), c_null_ptr)
        if (rc /= 0) call fail('mct_fm2d_times', rc)
    end subroutine

    ! The same for group-velocity data (settings%phaseGroup == 1, uar = 0): travel times, the rays (phaseRays(:, period), stored
    ! in the slot dat%raystat(:,2,period) names, as rpaths does) and crazyray(period).  The caller goes on exactly as after
    ! the Fortran modrays: like%srdist = phaseRays%length(), `any(crazyray > 0)` -> huge, CalGroupTime(like%gvel, ...).
    subroutine fm2d_rays_b200(src, rev, raystat, grid, vel, gridx, gridy, sgref, sgdic, sgext, order, band, phaseTime, phaseRays, crazyray)
        use m_fm2d, only : T_RAY
        real(c_double), dimension(:,:), intent(in) :: src, rev
        integer(c_int), dimension(:,:,:), intent(in) :: raystat
        type(T_GRID), intent(in) :: grid
        real(c_double), dimension(:,:,:), intent(in) :: vel
        integer, intent(in) :: gridx, gridy, sgref, sgdic, sgext, order
        real(c_double), intent(in) :: band
        real(c_double), dimension(:,:,:), intent(inout), target :: phaseTime
        type(T_RAY), dimension(:,:), intent(inout) :: phaseRays          ! (nrev*nsrc, np)
        integer, dimension(:), intent(inout) :: crazyray

        real(c_double), allocatable, target :: sx(:), sz(:), rx(:), rz(:), maps(:,:,:), rpts(:,:,:,:), rlen(:,:)
        integer(c_int), allocatable, target :: srs(:,:), srsv(:,:), rn(:,:), cz(:)
        type(mct_fm2d_opts) :: o
        integer(c_int) :: rc, cap
        integer :: i, n, np, nsrc, nrev, nrr

        nsrc = size(src, 2); nrev = size(rev, 2); np = size(vel, 1); nrr = nrev*nsrc
        cap = 8 * ((grid%nx - 1)*gridx + 1 + (grid%ny - 1)*gridy + 1)
        allocate(sx(nsrc), sz(nsrc), rx(nrev), rz(nrev), srs(nrr, np), srsv(nrr, np), maps(size(vel, 2), size(vel, 3), np))
        allocate(rn(nrr, np), rpts(2, cap, nrr, np), rlen(nrr, np), cz(np))
        sx = src(1,:); sz = src(2,:); rx = rev(1,:); rz = rev(2,:)
        srs = raystat(:, 1, :); srsv = raystat(:, 2, :)
        do i = 1, np
            maps(:, :, i) = vel(i, :, :)
        enddo
        o%gridx = gridx; o%gridy = gridy; o%sgref = sgref; o%sgdic = sgdic; o%sgext = sgext; o%order = order; o%band = band
        rc = mct_fm2d_rays(c_loc(sx), c_loc(sz), int(nsrc, c_int), c_loc(rx), c_loc(rz), int(nrev, c_int), c_loc(srs), c_loc(srsv), &
                           c_loc(maps), int(np, c_int), int(grid%nx, c_int), int(grid%ny, c_int), grid%xmin, grid%ymin, grid%dx, grid%dy, &
                           o, c_loc(phaseTime), cap, c_loc(rn), c_loc(rpts), c_loc(rlen), c_loc(cz))
        if (rc /= 0) call fail('mct_fm2d_rays', rc)
        do i = 1, np
            crazyray(i) = crazyray(i) + cz(i)
            do n = 1, nrr
                if (rn(n, i) < 1) cycle
                if (allocated(phaseRays(n, i)%points)) deallocate(phaseRays(n, i)%points)
                allocate(phaseRays(n, i)%points(2, rn(n, i)))
                phaseRays(n, i)%points = rpts(:, 1:rn(n, i), n, i)
                phaseRays(n, i)%npoints = rn(n, i)
            enddo
        enddo
    end subroutine

    ! The ierr = 2 columns of a dispersion call, through the reference's own surfmodes (GRT branch).
    subroutine solve_lvl_columns_on_host(model, grid, ix0, ix1, iy0, iy1, freqs, raylov, phaseGroup, dPhaseVel, var, pvel, gvel, ierr, tol)
        use m_surfmodes, only : surfmodes, T_MODES_PARA
        type(T_MOD), intent(in) :: model
        type(T_GRID), intent(in) :: grid
        integer, intent(in) :: ix0, ix1, iy0, iy1, raylov, phaseGroup, var
        real(c_double), dimension(:), intent(in) :: freqs
        real(c_double), intent(in) :: dPhaseVel
        real(c_double), dimension(:,:,:), intent(inout) :: pvel, gvel
        integer(c_int), dimension(:,:), intent(inout) :: ierr
        real(c_double), intent(in), optional :: tol          ! settings%tol (GRT only); default as examples/example1/MCTomo.inp
        type(T_MODES_PARA) :: paras
        real(c_double), allocatable :: thick(:), vp(:), vs(:), rho(:)
        integer :: i, j, n, e
        ! src/likelihood_surf.F90:173-182
        paras%modetype = raylov
        paras%phaseGroup = phaseGroup
        paras%tolmin = 1.0E-6_c_double
        if (present(tol)) paras%tolmin = tol
        paras%tolmax = 10 * paras%tolmin
        paras%smin_min = 1E-3
        paras%smin_max = 5E-3
        paras%dc = dPhaseVel
        paras%dcm = dPhaseVel
        paras%dc1 = dPhaseVel
        paras%dc2 = dPhaseVel
        do i = ix0, ix1
            do j = iy0, iy1
                if (ierr(j - iy0 + 1, i - ix0 + 1) /= 2) cycle
                call column_to_layers(model, grid, i, j, var, thick, vp, vs, rho, n)
                call surfmodes(thick(1:n), vp(1:n), vs(1:n), rho(1:n), freqs, paras, pvel(:, j - iy0 + 1, i - ix0 + 1), &
                               gvel(:, j - iy0 + 1, i - ix0 + 1), e)
                ierr(j - iy0 + 1, i - ix0 + 1) = e
            enddo
        enddo
    end subroutine

    ! convert_to_layer for ONE column (src/likelihood_surf.F90:541-625 / forward_modelling.f90:72-175, var selects EPS
    ! and the /scaling step); only used on the rare GRT columns above.
    subroutine column_to_layers(model, grid, i, j, var, thick, vp, vs, rho, n)
        type(T_MOD), intent(in) :: model
        type(T_GRID), intent(in) :: grid
        integer, intent(in) :: i, j, var
        real(c_double), allocatable, intent(out) :: thick(:), vp(:), vs(:), rho(:)
        integer, intent(out) :: n
        real(c_double) :: eps, last_vs
        integer :: k, last_k
        eps = real(1.0E-10, c_double)
        if (var /= 0) eps = real(1.0E-5, c_double)
        allocate(thick(grid%nz + 2), vp(grid%nz + 2), vs(grid%nz + 2), rho(grid%nz + 2))
        n = 0
        if (grid%waterDepth > merge(eps, 0.0_c_double, var == 0)) then
            n = 1
            thick(1) = grid%waterDepth; vp(1) = 1.5_c_double; vs(1) = 0.0_c_double; rho(1) = 1.0_c_double
        endif
        last_vs = model%vs(1, j, i); last_k = 1
        do k = 2, grid%nz
            if (abs(model%vs(k, j, i) - last_vs) > eps) then
                n = n + 1
                thick(n) = (k - last_k) * grid%dz
                vp(n) = model%vp(last_k, j, i); vs(n) = model%vs(last_k, j, i); rho(n) = model%rho(last_k, j, i)
                last_vs = model%vs(k, j, i); last_k = k
            endif
        enddo
        n = n + 1                                            ! the last run is the half-space (thickness 0)
        thick(n) = 0.0_c_double
        vp(n) = model%vp(grid%nz, j, i); vs(n) = model%vs(grid%nz, j, i); rho(n) = model%rho(grid%nz, j, i)
        if (var == 0) then
            thick(1:n) = thick(1:n) / grid%scaling
            if (grid%waterDepth > eps) thick(1) = grid%waterDepth
        endif
    end subroutine

    ! ---- resident session: the sampler's per-iteration sequence (src/mcmc_loc2.f90:199-228,556-566) with the chain's
    !      model and maps kept on the device.  One session per chain; calls mirror the sampler's own steps:
    !        mcmc :139-143          -> b200_session_set_model      (kdtree_to_grid(full) + likelihood's dispersion)
    !        case birth/death/move/value :220-228,279-282,328-330,395-397
    !                                -> b200_session_propose        (kdtree_to_grid(box[,pm]) + dispersion of box + halo)
    !        surf_likelihood's tail :226-243,356-404 (straight rays) -> b200_session_likelihood
    !        accept :244-265 / reject :554-567 -> b200_session_accept / b200_session_reject
    !        every `thin` samples :587-600  -> b200_session_stat_rti (kdtree_to_grid(full) is not needed: the model IS resident)
    subroutine b200_session_create(S, grid, freqs, raylov, phaseGroup, dPhaseVel)
        type(T_B200_SESSION), intent(out) :: S
        type(T_GRID), intent(in) :: grid
        real(c_double), dimension(:), intent(in), target :: freqs
        integer, intent(in) :: raylov, phaseGroup
        real(c_double), intent(in) :: dPhaseVel
        type(mct_disp_opts) :: opt
        integer(c_int) :: rc
        opt%raylov = raylov; opt%phaseGroup = phaseGroup; opt%nmodes = 0; opt%check_scope = 0
        opt%dphase = dPhaseVel
        opt%layer_eps = real(1.0E-10, c_double); opt%water_thresh = real(1.0E-10, c_double); opt%preset = 100.0_c_double
        rc = mct_session_create(c_grid(grid), c_loc(freqs), size(freqs), opt, 1_c_int, S%h)
        if (rc /= 0) call fail('mct_session_create', rc)
        S%nx = grid%nx; S%ny = grid%ny; S%nout = size(freqs)
        allocate(S%pvel_w(S%nout * grid%nx * grid%ny), S%gvel_w(S%nout * grid%nx * grid%ny), S%ierr_w(grid%nx * grid%ny))
    end subroutine

    subroutine b200_session_destroy(S)
        type(T_B200_SESSION), intent(inout) :: S
        integer(c_int) :: rc
        rc = mct_session_destroy(S%h)
        S%h = c_null_ptr
    end subroutine

    ! full evaluation; pvel/gvel (np,ny,nx), ierr (ny,nx) as surf_likelihood holds them; invalid = check_model's answer
    subroutine b200_session_set_model(S, RTI, pvel, gvel, ierr, invalid)
        type(T_B200_SESSION), intent(inout) :: S
        type(T_RUN_INFO), intent(in), target :: RTI
        real(c_double), dimension(:,:,:), intent(inout), target :: pvel, gvel
        integer(c_int), dimension(:,:), intent(inout), target :: ierr
        logical, intent(out) :: invalid
        integer(c_int), target :: inval
        integer(c_int) :: rc
        inval = 0
        rc = mct_session_set_model(S%h, c_loc(RTI%points), c_loc(RTI%parameters), int(RTI%ncells, c_int), c_loc(pvel), c_loc(gvel), &
                                   c_loc(ierr), c_loc(inval))
        invalid = (inval /= 0)
        if (rc /= 0 .and. rc /= 2) call fail('mct_session_set_model', rc)
    end subroutine

    ! one proposal: RTI holds the nuclei AFTER the move, bnd_box what the sampler hands to kdtree_to_grid, pm the moved
    ! cell's old (vp,vs,rho) for a value move.  On return S%win = (ix0,ix1,iy0,iy1) and S%pvel_w/gvel_w/ierr_w hold the
    ! window's maps packed (np, iy0:iy1, ix0:ix1): copy them into like%vel / like%gvel exactly as
    ! likelihood_surf.F90:226-231,259-264 does with its local pvel/gvel.  bad = any(ierr == 1) of the window.
    subroutine b200_session_propose(S, RTI, bnd_box, invalid, bad, pm)
        type(T_B200_SESSION), intent(inout), target :: S
        type(T_RUN_INFO), intent(in), target :: RTI
        type(d3), dimension(2), intent(in) :: bnd_box
        logical, intent(out) :: invalid, bad
        type(p3), intent(in), optional :: pm
        real(c_double) :: box(6)
        real(c_double), target :: pmv(3)
        type(c_ptr) :: pm_ptr
        integer(c_int) :: rc, inval
        integer :: ncol
        box = [bnd_box(1)%x, bnd_box(1)%y, bnd_box(1)%z, bnd_box(2)%x, bnd_box(2)%y, bnd_box(2)%z]
        pm_ptr = c_null_ptr
        if (present(pm)) then
            pmv = [pm%vp, pm%vs, pm%rho]
            pm_ptr = c_loc(pmv)
        endif
        rc = mct_session_propose(S%h, c_loc(RTI%points), c_loc(RTI%parameters), int(RTI%ncells, c_int), box, pm_ptr, S%win, &
                                 c_loc(S%pvel_w), c_loc(S%gvel_w), c_loc(S%ierr_w), inval)
        if (rc /= 0 .and. rc /= 2) call fail('mct_session_propose', rc)
        invalid = (inval /= 0)
        ncol = max(0, S%win(2) - S%win(1) + 1) * max(0, S%win(4) - S%win(3) + 1)
        bad = .false.
        if (.not. invalid .and. ncol > 0) bad = any(S%ierr_w(1:ncol) /= 0)
    end subroutine

    subroutine b200_session_accept(S)
        type(T_B200_SESSION), intent(inout) :: S
        integer(c_int) :: rc
        rc = mct_session_accept(S%h)
        if (rc /= 0) call fail('mct_session_accept', rc)
    end subroutine

    subroutine b200_session_reject(S)
        type(T_B200_SESSION), intent(inout) :: S
        integer(c_int) :: rc
        rc = mct_session_reject(S%h)
        if (rc /= 0) call fail('mct_session_reject', rc)
    end subroutine

    ! surf_likelihood's tail for straight rays (likelihood_surf.F90:226-243,356-404) on the resident maps.  The rays
    ! (setup_straightRays, once) and the data (dat%ttime, dat%raystat, like%srdist) are made resident by
    ! mct_session_set_rays / mct_session_set_data at start-up; per call only the noise parameters travel.
    subroutine b200_session_likelihood(S, pending, snoise0, snoise1, like_val, misfit, unweighted_misfit)
        type(T_B200_SESSION), intent(inout) :: S
        logical, intent(in) :: pending
        real(c_double), dimension(:), intent(in), target :: snoise0, snoise1
        real(c_double), intent(out) :: like_val, misfit, unweighted_misfit
        real(c_double) :: out(3)
        integer(c_int) :: rc
        rc = mct_session_likelihood(S%h, merge(1_c_int, 0_c_int, pending), c_null_ptr, c_null_ptr, 0_c_int, c_loc(snoise0), &
                                    c_loc(snoise1), out, c_null_ptr, c_null_ptr)
        if (rc /= 0) call fail('mct_session_likelihood', rc)      ! rc = 6: 'The noise level is 0!' (likelihood_surf.F90:387-390)
        like_val = out(1); misfit = out(2); unweighted_misfit = out(3)
    end subroutine

    ! stat_rti's four sums (src/mcmc_loc2.f90:1966-1978) on the resident model; read back with mct_session_stat_get
    subroutine b200_session_stat_rti(S)
        type(T_B200_SESSION), intent(inout) :: S
        integer(c_int) :: rc
        rc = mct_session_stat_accumulate(S%h)
        if (rc /= 0) call fail('mct_session_stat_accumulate', rc)
    end subroutine

    subroutine fail(what, rc)
        use m_exception, only : exception_raiseError
        character(len=*), intent(in) :: what
        integer(c_int), intent(in) :: rc
        character(len=512), pointer :: msg
        character(len=16) :: code
        call c_f_pointer(mct_last_error(), msg)
        write(code, '(I0)') rc
        call exception_raiseError(what // ' failed with status ' // trim(code) // ': ' // msg(1:index(msg, c_null_char) - 1))
    end subroutine

end module m_mctomo_b200
