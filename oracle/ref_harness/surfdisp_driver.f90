! surfdisp_driver.f90 -- test harness around the REFERENCE's surfdisp96 / surfdisp_mmodes (surfmodes/surfdisp96.f),
! built by oracle/build_ref_surfdisp.sh where a Fortran compiler exists.  TEST INFRASTRUCTURE.
!
! stdin (list-directed text, one case after another until end of file):
!     nlayer iwave mode igr kmax mmode dphase        mmode = 0: surfdisp96, 1: surfdisp_mmodes
!     thk(1:nlayer)   vp(1:nlayer)   vs(1:nlayer)   rho(1:nlayer)     (real*4, as surfmodes.f90:81-83 narrows them)
!     t(1:kmax)                                                       (double precision periods)
! stdout per case: ierr, then cp(1:kmax*mode) and cg(1:kmax*mode) as hexadecimal bit patterns (Z16.16), mode-major
! (surfmodes.f90:179-180), so the comparison in tests/test_oracle_vs_fortran.py is bit for bit.
program surfdisp_driver
    implicit none
    integer, parameter :: NLAY = 200, NP = 60
    real(kind=4) :: thk(NLAY), vp(NLAY), vs(NLAY), rho(NLAY)
    double precision :: t(NP), dphase
    double precision, allocatable :: cp(:), cg(:), cp2(:,:), cg2(:,:)
    integer :: nlayer, iwave, mode, igr, kmax, mmode, ierr, ios, i, j
    do
        read(*, *, iostat=ios) nlayer, iwave, mode, igr, kmax, mmode, dphase
        if (ios /= 0) exit
        thk = 0; vp = 0; vs = 0; rho = 0; t = 0
        read(*, *) thk(1:nlayer)
        read(*, *) vp(1:nlayer)
        read(*, *) vs(1:nlayer)
        read(*, *) rho(1:nlayer)
        read(*, *) t(1:kmax)
        if (mmode == 0) then
            allocate(cp(kmax), cg(kmax))
            call surfdisp96(thk, vp, vs, rho, nlayer, 0, iwave, 1, igr, kmax, t, dphase, cp, cg, ierr)
            write(*, '(I4)') ierr
            write(*, '(4(1X,Z16.16))') (cp(i), i = 1, kmax)
            write(*, '(4(1X,Z16.16))') (cg(i), i = 1, kmax)
            deallocate(cp, cg)
        else
            allocate(cp2(kmax, mode), cg2(kmax, mode))
            call surfdisp_mmodes(thk, vp, vs, rho, nlayer, 0, iwave, mode, igr, kmax, t, dphase, cp2, cg2, ierr)
            write(*, '(I4)') ierr
            write(*, '(4(1X,Z16.16))') ((cp2(i, j), i = 1, kmax), j = 1, mode)
            write(*, '(4(1X,Z16.16))') ((cg2(i, j), i = 1, kmax), j = 1, mode)
            deallocate(cp2, cg2)
        end if
    end do
end program surfdisp_driver
