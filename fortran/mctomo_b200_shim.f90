! mctomo_b200_shim.f90 -- ISO_C_BINDING glue between the unchanged MCTomo host and libmctomo_b200.so.
!
! This file is SOURCE ONLY: the build image of this repository has no Fortran compiler, so it has not
! been compiled here.  It is written against the reference's own types and mirrors the style of the
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

    ! mirrors `mct_grid` of include/mctomo_b200.h
    type, bind(C) :: mct_grid
        integer(c_int32_t) :: nx, ny, nz
        real(c_double)     :: xmin, ymin, zmin
        real(c_double)     :: dx, dy, dz
        real(c_double)     :: waterDepth
        real(c_double)     :: scaling
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
    end interface

contains

    subroutine mctomo_b200_init(rank)
        integer, intent(in) :: rank
        integer(c_int) :: rc
        ! chains are MPI ranks (src/MCTomo.F90:131-133); eight GPUs per node, one or more ranks per GPU
        rc = mct_init(int(mod(rank, 8), c_int))
        if (rc /= 0) call fail('mct_init', rc)
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
                                    dPhaseVel, pvel, gvel, ierr, invalid, variant)
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

        type(mct_grid) :: g
        type(mct_disp_opts) :: opt
        integer(c_int), target :: inval
        integer(c_int) :: rc
        integer :: var

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
        inval = 0
        if (var == 0) then
            rc = mct_surf_dispersion(c_loc(model%vp), c_loc(model%vs), c_loc(model%rho), g, ix0, ix1, iy0, iy1, &
                                     c_loc(freqs), size(freqs), opt, c_loc(pvel), c_loc(gvel), c_loc(ierr), c_loc(inval))
        else
            rc = mct_surf_dispersion(c_loc(model%vp), c_loc(model%vs), c_loc(model%rho), g, ix0, ix1, iy0, iy1, &
                                     c_loc(freqs), size(freqs), opt, c_loc(pvel), c_loc(gvel), c_loc(ierr), c_null_ptr)
        endif
        invalid = (inval /= 0)
        ! rc = 2: some column has a low-velocity layer and needs the generalized R/T branch
        ! (surfmodes.f90:84-87); those columns carry ierr = 2 and the caller may run the Fortran surfmodes on them.
        if (rc < 0 .or. rc == 1) call fail('mct_surf_dispersion', rc)
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
