! pic1dp_refdump.F90 -- NOT part of the reference and NOT part of the product: a dump module that
! oracle/ref_recipe/make_ref_dump.sh drops into a COPY of the reference source tree so that the reference
! itself (Fortran 2003 + PETSc + MPI) writes its per-substep hot-path state.  The dump pins the CPU oracle
! (oracle/pic1dp_oracle.c) to the real program: tests/test_oracle_ref_pin.py replays it.
!
! Written against the module variables of /root/reference/src/pic1dp_particle.F90:26-58 and
! /root/reference/src/pic1dp_field.F90:27-48; access pattern (VecGetArrayF90 / VecRestoreArrayF90) as at
! /root/reference/src/pic1dp_interaction.F90:221-237.  UNTESTED in this repository's image (no gfortran / PETSc).
!
! Record layout of refdump_rank<mype>.bin (stream access, native endianness, appended once per call):
!   int32  magic = 20140512, itime, irk, npe, mype, nx, nmode, nspecies
!   per species: int64 np; real64 x(np), v(np), p(np), w(np)
!   int32  ix_low, ix_high; real64 chargeden(ix_low:ix_high-1), electric(ix_low:ix_high-1)
!   int32  imode_low, imode_high; real64 mode_re(imode_low:imode_high-1), mode_im(imode_low:imode_high-1)
module pic1dp_refdump
use pic1dp_input
implicit none
#include "finclude/petscdef.h"

contains

subroutine refdump_state(itime, irk)
use pic1dp_global
use pic1dp_input
use pic1dp_particle
use pic1dp_field
implicit none
#include "finclude/petsc.h90"

PetscInt, intent(in) :: itime, irk

PetscScalar, dimension(:), pointer :: pa
integer :: u, ispecies, n
character(len = 64) :: fname

write (fname, '(a, i4.4, a)') 'refdump_rank', global_mype, '.bin'
open (newunit = u, file = trim(fname), access = 'stream', form = 'unformatted', &
  position = 'append', status = 'unknown')

write (u) int(20140512, 4), int(itime, 4), int(irk, 4), int(global_npe, 4), &
  int(global_mype, 4), int(input_nx, 4), int(input_nmode, 4), int(input_nspecies, 4)

do ispecies = 1, input_nspecies
  n = int(particle_np(ispecies))
  write (u) int(n, 8)
  call VecGetArrayF90(particle_x(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  write (u) real(pa(1 : n), 8)
  call VecRestoreArrayF90(particle_x(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  call VecGetArrayF90(particle_v(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  write (u) real(pa(1 : n), 8)
  call VecRestoreArrayF90(particle_v(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  call VecGetArrayF90(particle_p(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  write (u) real(pa(1 : n), 8)
  call VecRestoreArrayF90(particle_p(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  call VecGetArrayF90(particle_w(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
  write (u) real(pa(1 : n), 8)
  call VecRestoreArrayF90(particle_w(ispecies), pa, global_ierr)
  CHKERRQ(global_ierr)
end do

n = int(field_ix_high - field_ix_low)
write (u) int(field_ix_low, 4), int(field_ix_high, 4)
call VecGetArrayF90(field_chargeden, pa, global_ierr)
CHKERRQ(global_ierr)
write (u) real(pa(1 : n), 8)
call VecRestoreArrayF90(field_chargeden, pa, global_ierr)
CHKERRQ(global_ierr)
call VecGetArrayF90(field_electric, pa, global_ierr)
CHKERRQ(global_ierr)
write (u) real(pa(1 : n), 8)
call VecRestoreArrayF90(field_electric, pa, global_ierr)
CHKERRQ(global_ierr)

n = int(field_imode_high - field_imode_low)
write (u) int(field_imode_low, 4), int(field_imode_high, 4)
call VecGetArrayF90(field_mode_re, pa, global_ierr)
CHKERRQ(global_ierr)
write (u) real(pa(1 : n), 8)
call VecRestoreArrayF90(field_mode_re, pa, global_ierr)
CHKERRQ(global_ierr)
call VecGetArrayF90(field_mode_im, pa, global_ierr)
CHKERRQ(global_ierr)
write (u) real(pa(1 : n), 8)
call VecRestoreArrayF90(field_mode_im, pa, global_ierr)
CHKERRQ(global_ierr)

close (u)

end subroutine refdump_state

end module pic1dp_refdump
