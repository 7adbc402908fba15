!
!  greb_b200_host.f90 -- ISO_C_BINDING layer between the reference's Fortran host and the
!  B200-native stepping core (include/greb_b200.h, libgreb_b200.so).
!
!  STATUS: written against the reference interfaces but NOT compiled here -- there is no Fortran
!  compiler in the build image (gfortran / flang / nvfortran absent).  The same host logic is
!  implemented and tested in Python over the same C ABI (greb_b200/host.py: run_namelists);
!  this file is what a maintainer of sieste/greb-climate-model adds.
!
!  How to use it with the reference:
!     1. keep src/greb.f90 modules mo_numerics, mo_physics, mo_diagnostics and PROGRAM greb_run
!        (namelist, input/ files, co2 padding, Toclim) unchanged;
!     2. replace the body of `subroutine greb_model` (src/greb.f90:161-236) by the body of
!        greb_model_b200 below (or rename and call it from greb_run at :1096);
!     3. gfortran -O3 src/greb.f90 greb_b200_host.f90 -L. -lgreb_b200 -o greb
!
!  The subroutine interfaces, namelist, input files and the output/scenario record layout
!  (src/greb.f90:978-982: Tsurf, Tair, Tocean, q, albedo per month) are the reference's own.
!
module greb_b200_c
  use, intrinsic :: iso_c_binding
  implicit none

  ! struct greb_physics_par (include/greb_b200.h) == namelist physics_par + co2_flux
  type, bind(C) :: greb_physics_par
     real(c_float) :: pi, sig, rho_ocean, rho_land, rho_air, cp_ocean, cp_land, cp_air, eps
     real(c_float) :: d_ocean, d_land, d_air, ct_sens, da_ice, a_no_ice, a_cloud
     real(c_float) :: Tl_ice1, Tl_ice2, To_ice1, To_ice2, co_turb, kappa, ce, cq_latent, cq_rain
     real(c_float) :: z_air, z_vapor, r_qviwv
     real(c_float) :: p_emi(10)
     real(c_float) :: co2_flux
  end type greb_physics_par

  interface
     integer(c_int) function greb_b200_create(h, n_members, device) bind(C, name='greb_b200_create')
       import :: c_ptr, c_int
       type(c_ptr), intent(out) :: h
       integer(c_int), value :: n_members, device
     end function
     integer(c_int) function greb_b200_destroy(h) bind(C, name='greb_b200_destroy')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     type(c_ptr) function greb_b200_last_error(h) bind(C, name='greb_b200_last_error')
       import :: c_ptr
       type(c_ptr), value :: h
     end function
     integer(c_int) function greb_b200_set_forcing(h, z_topo, glacier, sw_solar, tclim, qclim, swetclim, &
          uclim, vclim, mldclim, cldclim) bind(C, name='greb_b200_set_forcing')
       import :: c_ptr, c_int, c_float
       type(c_ptr), value :: h
       real(c_float), intent(in) :: z_topo(*), glacier(*), sw_solar(*), tclim(*), qclim(*), swetclim(*), &
            uclim(*), vclim(*), mldclim(*), cldclim(*)
     end function
     integer(c_int) function greb_b200_set_member(h, member, p, co2_ppm, n_years, year0) &
          bind(C, name='greb_b200_set_member')
       import :: c_ptr, c_int, c_float, greb_physics_par
       type(c_ptr), value :: h
       integer(c_int), value :: member, n_years, year0
       type(greb_physics_par), intent(in) :: p
       real(c_float), intent(in) :: co2_ppm(*)
     end function
     ! process switches = the log_exp experiments of greb.original.model.f90 (GREB_SW_* bit mask)
     integer(c_int) function greb_b200_set_switches(h, member, mask) bind(C, name='greb_b200_set_switches')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: member, mask
     end function
     integer(c_int) function greb_b200_init(h) bind(C, name='greb_b200_init')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function greb_b200_spinup(h, years) bind(C, name='greb_b200_spinup')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: years
     end function
     integer(c_int) function greb_b200_reset_scenario(h) bind(C, name='greb_b200_reset_scenario')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function greb_b200_run(h, years, out, out_members, n_out, gmean, gmean_coslat) &
          bind(C, name='greb_b200_run')
       import :: c_ptr, c_int, c_float
       type(c_ptr), value :: h
       integer(c_int), value :: years, n_out
       real(c_float), intent(out) :: out(*)
       type(c_ptr), value :: out_members        ! NULL = all members
       real(c_float), intent(out) :: gmean(*), gmean_coslat(*)
     end function
     integer(c_int) function greb_b200_get_state(h, member, which, out) bind(C, name='greb_b200_get_state')
       import :: c_ptr, c_int, c_float
       type(c_ptr), value :: h
       integer(c_int), value :: member, which
       real(c_float), intent(out) :: out(*)
     end function
  end interface

contains

  subroutine greb_b200_check(h, rc, what)
    type(c_ptr), intent(in) :: h
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc == 0) return
    call c_f_pointer(greb_b200_last_error(h), msg, [512])
    n = 1
    do while (n < 512 .and. msg(n) /= c_null_char)
       n = n + 1
    end do
    print *, 'greb_b200: ', what, ' failed, rc = ', rc, ': ', msg(1:n-1)
    stop 1
  end subroutine

end module greb_b200_c


!+++++++++++++++++++++++++++++++++++++++
subroutine greb_model_b200
!+++++++++++++++++++++++++++++++++++++++
!   drop-in replacement for the body of greb_model (src/greb.f90:161-236): same module inputs
!   (mo_numerics, mo_physics, mo_diagnostics), same output file, single member.

  use mo_numerics
  use mo_physics
  use mo_diagnostics
  use greb_b200_c
  implicit none

  type(c_ptr) :: h
  type(greb_physics_par) :: p
  real(c_float), allocatable :: out(:,:,:,:,:), gmean(:), gmean_w(:)
  integer :: irec, iy, im, iv

  call greb_b200_check(c_null_ptr, greb_b200_create(h, 1_c_int, 0_c_int), 'create')

  ! the ten inputs PROGRAM greb_run has read (src/greb.f90:1073-1085); Fortran (i,j,n) storage is
  ! exactly the C [n][j][i] layout the library expects -- no transposes
  call greb_b200_check(h, greb_b200_set_forcing(h, z_topo, glacier, sw_solar, Tclim, qclim, swetclim, &
       uclim, vclim, mldclim, cldclim), 'set_forcing')

  p%pi = pi; p%sig = sig; p%rho_ocean = rho_ocean; p%rho_land = rho_land; p%rho_air = rho_air
  p%cp_ocean = cp_ocean; p%cp_land = cp_land; p%cp_air = cp_air; p%eps = eps
  p%d_ocean = d_ocean; p%d_land = d_land; p%d_air = d_air; p%ct_sens = ct_sens; p%da_ice = da_ice
  p%a_no_ice = a_no_ice; p%a_cloud = a_cloud; p%Tl_ice1 = Tl_ice1; p%Tl_ice2 = Tl_ice2
  p%To_ice1 = To_ice1; p%To_ice2 = To_ice2; p%co_turb = co_turb; p%kappa = kappa; p%ce = ce
  p%cq_latent = cq_latent; p%cq_rain = cq_rain; p%z_air = z_air; p%z_vapor = z_vapor
  p%r_qviwv = r_qviwv; p%p_emi = p_emi; p%co2_flux = co2_flux

  ! co2_ppm(1:time_scnr) is already padded by greb_run (src/greb.f90:1053-1061)
  call greb_b200_check(h, greb_b200_set_member(h, 0_c_int, p, co2_ppm, int(time_scnr, c_int), &
       int(year0, c_int)), 'set_member')
  call greb_b200_check(h, greb_b200_init(h), 'init')

  print*,'% FLUX CORRECTION RUN; years = ', time_flux, ' co2 = ', CO2_flux
  call greb_b200_check(h, greb_b200_spinup(h, int(time_flux, c_int)), 'spinup')

  print*,'% MODEL RUN; years = ', time_scnr
  print*,'% saving output in file ', output_file_full
  allocate(out(xdim, ydim, 5, 12, time_scnr), gmean(time_scnr), gmean_w(time_scnr))
  call greb_b200_check(h, greb_b200_reset_scenario(h), 'reset_scenario')
  call greb_b200_check(h, greb_b200_run(h, int(time_scnr, c_int), out, c_null_ptr, 1_c_int, gmean, gmean_w), 'run')

  ! the reference's record stream on unit 22 (src/greb.f90:174, 978-982)
  open(22, file=output_file_full, ACCESS='DIRECT', FORM='UNFORMATTED', RECL=ireal*xdim*ydim)
  irec = 0
  print *, 'console output: year, co2, global avg temp, avg temp for ipx/ipy'
  do iy = 1, time_scnr
     do im = 1, 12
        do iv = 1, 5
           irec = irec + 1
           write(22, rec=irec) out(:, :, iv, im, iy)
        end do
     end do
     print *, real(year0 + iy - 1), co2_ppm(iy), gmean(iy)
  end do
  close(22)

  call greb_b200_check(h, greb_b200_destroy(h), 'destroy')

end subroutine greb_model_b200
