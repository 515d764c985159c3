! ISO_C_BINDING interface to libsbdart_b200.so (include/sbdart_b200.h) for a
! Fortran-2003 host such as SBDART's drt.f.  Mirrors the C structs field by field.
! Not compiled in this repository's image (no Fortran compiler is installed);
! INTEGRATION.md shows where the calls go in drt.f:529-560.
module sbd_b200
  use iso_c_binding
  implicit none

  type, bind(c) :: sbd_dims
     integer(c_int32_t) :: nbins, nlyr, nstr, nmom, ntau, numu, nphi, ncol
  end type sbd_dims

  type, bind(c) :: sbd_bin
     real(c_double) :: fbeam, umu0, phi0, fisot, albedo, btemp, ttemp, temis, wvnmlo, wvnmhi, accur
     integer(c_int32_t) :: plank, col
  end type sbd_bin

  integer(c_int), parameter :: SBD_SUCCESS = 0, SBD_ERR_CUDA = -100, SBD_ERR_ARG = -101, &
                               SBD_ERR_UNSUPPORTED = -102
  ! per-bin status (errmsg numbers of the reference, disutil.f:280-330)
  integer(c_int32_t), parameter :: SBD_BIN_OK = 0, SBD_BIN_ANGLE_CLASH = 1, SBD_BIN_BAD_INPUT = -1, &
                                   SBD_BIN_EIG_FAIL = -2, SBD_BIN_SINGULAR = -3

  interface
     integer(c_int) function sbd_create(h, device) bind(c, name='sbd_create')
       import
       type(c_ptr), intent(out) :: h
       integer(c_int), value :: device
     end function sbd_create

     subroutine sbd_destroy(h) bind(c, name='sbd_destroy')
       import
       type(c_ptr), value :: h
     end subroutine sbd_destroy

     ! host buffers in, host buffers out; H2D, kernel and D2H are pipelined inside
     integer(c_int) function sbd_disort_batch(h, dims, dtauc, ssalb, pmom, bins, temper, utau, &
          umu, phi, rfldir, rfldn, flup, dfdt, uavg, uu, status) bind(c, name='sbd_disort_batch')
       import
       type(c_ptr), value :: h
       type(sbd_dims), intent(in) :: dims
       real(c_double), intent(in) :: dtauc(*), ssalb(*), pmom(*), temper(*)
       type(sbd_bin), intent(in) :: bins(*)
       type(c_ptr), value :: utau, umu, phi, uu        ! c_null_ptr for flux runs
       real(c_double), intent(out) :: rfldir(*), rfldn(*), flup(*), dfdt(*), uavg(*)
       integer(c_int32_t), intent(out) :: status(*)
     end function sbd_disort_batch

     ! CORINT (disort.f:112-118): Nakajima-Tanaka corrections on the following radiance calls
     integer(c_int) function sbd_set_corint(h, on) bind(c, name='sbd_set_corint')
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value :: h
       integer(c_int32_t), value :: on
     end function sbd_set_corint

     ! intensities only at these output levels (ntop / nbot of drt.f:1008-1016)
     integer(c_int) function sbd_set_radiance_levels(h, levels, n) bind(c, name='sbd_set_radiance_levels')
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value :: h
       integer(c_int32_t), intent(in) :: levels(*)
       integer(c_int32_t), value :: n
     end function sbd_set_radiance_levels

     ! BRDF surfaces (LAMBER = .FALSE., isalb 7-9): SURFAC's tables (disort.f:3765-3907) of nsurf
     ! surfaces, bdr(0:n, n, nmodes, nsurf), bem(n, nsurf), rmu(0:n, numu, nmodes, nsurf),
     ! emu(numu, nsurf) in Fortran order; bins(b)%albedo = -(s+1) selects surface s (0-based)
     integer(c_int) function sbd_set_surfaces(h, nsurf, nstr, nmodes, numu, bdr, bem, rmu, emu) &
          bind(c, name='sbd_set_surfaces')
       import :: c_ptr, c_int, c_int32_t, c_double
       type(c_ptr), value :: h
       integer(c_int32_t), value :: nsurf, nstr, nmodes, numu
       real(c_double), intent(in) :: bdr(*), bem(*)
       type(c_ptr), value :: rmu, emu                    ! c_null_ptr for flux runs
     end function sbd_set_surfaces

     ! the BDREF the single-call entry disort_ uses for LAMBER = .FALSE. (spectra.f:249);
     ! not needed when the executable exports its own bdref_ (link with -rdynamic)
     subroutine sbd_set_bdref_callback(f) bind(c, name='sbd_set_bdref_callback')
       import :: c_funptr
       type(c_funptr), value :: f
     end subroutine sbd_set_bdref_callback

     integer(c_int) function sbd_synchronize(h) bind(c, name='sbd_synchronize')
       import
       type(c_ptr), value :: h
     end function sbd_synchronize
  end interface

contains

  ! One batched solve of nb bins that were collected by the wavelength loop
  ! (drt.f:425-561): dtauc(nz,nb), ssalb(nz,nb), pmom(0:nmom,nz,nb) are exactly the
  ! C layouts [B][L] and [B][L][nmom+1].
  subroutine sbd_solve_bins(h, nz, nstr, nmom, nb, dtauc, ssalb, pmom, bins, temper, &
                            rfldir, rfldn, flup, dfdt, uavg, status, ierr)
    type(c_ptr), intent(in) :: h
    integer, intent(in) :: nz, nstr, nmom, nb
    real(c_double), intent(in) :: dtauc(nz, nb), ssalb(nz, nb), pmom(0:nmom, nz, nb), temper(0:nz)
    type(sbd_bin), intent(in) :: bins(nb)
    real(c_double), intent(out) :: rfldir(nz + 1, nb), rfldn(nz + 1, nb), flup(nz + 1, nb), &
                                   dfdt(nz + 1, nb), uavg(nz + 1, nb)
    integer(c_int32_t), intent(out) :: status(nb)
    integer, intent(out) :: ierr
    type(sbd_dims) :: d
    d = sbd_dims(nb, nz, nstr, nmom, 0, 0, 0, 1)
    ierr = sbd_disort_batch(h, d, dtauc, ssalb, pmom, bins, temper, c_null_ptr, c_null_ptr, &
                            c_null_ptr, rfldir, rfldn, flup, dfdt, uavg, c_null_ptr, status)
  end subroutine sbd_solve_bins

end module sbd_b200
