#!/bin/bash
# make_ref_dump.sh -- ONE COMMAND that pins the CPU oracle to the reference program on a box that has
# gfortran + MPI + PETSc (3.1 .. 3.5 layout, real scalars; README.md:35-38 of the reference).
#
#   PETSC_DIR=/path/to/petsc [PETSC_ARCH=...] oracle/ref_recipe/make_ref_dump.sh [REFERENCE_DIR] [NRANKS] [NMARKERS] [NSTEPS]
#
# It never modifies REFERENCE_DIR (default /root/reference) and copies no reference source into the repository:
# the tree is copied to oracle/_ref/work/ (git-ignored), edited there with the sed commands below, built with the
# reference's own build/Makefile, run under mpiexec, and the dumps are converted to tests/golden/ref_hotpath.npz
# by dump_to_golden.py.  tests/test_oracle_ref_pin.py then replays the oracle against the reference's own numbers.
#
# Edits to the COPY (each guarded by a grep on the line it expects, so a different reference version fails loudly):
#   src/pic1dp_input.F90:113  input_nparticle_max = 6400000      -> NMARKERS   (small, so the fixture stays small)
#   src/pic1dp_input.F90:223  input_multirand_seed_type = 3      -> 1          (constant, rank-dependent seeds)
#   src/pic1dp_input.F90:35   input_time_max = 500.0_kpr         -> NSTEPS*dt
#   src/pic1dp_input.F90:233  input_multirand_selftest = .true.  (kept)
#   src/pic1dp.F90            `use pic1dp_refdump`; `call refdump_state(0, 0)` after the initial field solve (:72);
#                             `call refdump_state(global_itime + 1, global_irk)` after the field solve of each substep (:89)
#   build/Makefile            pic1dp_refdump.o added to the program's objects
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
REF=${1:-/root/reference}
NR=${2:-2}
NM=${3:-200000}
NS=${4:-5}
: "${PETSC_DIR:?set PETSC_DIR (and PETSC_ARCH) to a PETSc build with real double scalars}"
command -v "${MPIF90:-mpif90}" >/dev/null || { echo "no MPI Fortran compiler (mpif90)"; exit 2; }
WORK=$REPO/oracle/_ref/work
rm -rf "$WORK"; mkdir -p "$WORK"
cp -r "$REF/src" "$REF/build" "$WORK/"
cp "$HERE/pic1dp_refdump.F90" "$WORK/src/"
I=$WORK/src/pic1dp_input.F90
M=$WORK/src/pic1dp.F90
guard() { sed -n "$2p" "$1" | grep -q "$3" || { echo "unexpected content at $1:$2 (want '$3')"; exit 3; }; }
guard "$I" 113 'input_nparticle_max = 6400000'
guard "$I" 223 'input_multirand_seed_type = 3'
guard "$I" 35 'input_time_max = 500.0_kpr'
guard "$M" 72 'call field_solve_electric'
guard "$M" 89 'call field_solve_electric'
TMAX=$(python3 -c "print('%.6f' % ($NS * 0.05 - 0.001))")
sed -i "113s/6400000/$NM/" "$I"
sed -i "223s/= 3/= 1/" "$I"
sed -i "35s/500.0_kpr/${TMAX}_kpr/" "$I"
# insert the calls bottom-up so that the line numbers above stay valid
sed -i "89a\    call refdump_state(global_itime + 1, global_irk)" "$M"
sed -i "72a\call refdump_state(0, 0)" "$M"
LINE=$(grep -n '^use pic1dp_output' "$M" | head -1 | cut -d: -f1)
[ -n "$LINE" ] || LINE=$(grep -n '^use ' "$M" | tail -1 | cut -d: -f1)
sed -i "${LINE}a\use pic1dp_refdump" "$M"
MK=$WORK/build/Makefile
sed -i 's/^pic1dp : pic1dp.o /pic1dp : pic1dp.o pic1dp_refdump.o /' "$MK"
sed -i 's/^pic1dp.o : \$(SRCDIR)\/pic1dp.F90 /pic1dp.o : $(SRCDIR)\/pic1dp.F90 pic1dp_refdump.o /' "$MK"
printf '\npic1dp_refdump.o : $(SRCDIR)/pic1dp_refdump.F90 pic1dp_global.o pic1dp_input.o pic1dp_particle.o pic1dp_field.o\n\t$(MPIF90) $(FFLAGS) -c -o $@ $< $(FCPPFLAGS)\n' >> "$MK"
make -C "$WORK/build" FFLAGS="-O3 -ffp-contract=off"
mkdir -p "$REPO/oracle/_ref/dump"; rm -f "$REPO/oracle/_ref/dump"/refdump_rank*.bin
(cd "$REPO/oracle/_ref/dump" && "${MPIEXEC:-mpiexec}" -n "$NR" "$WORK/build/pic1dp")
python3 "$HERE/dump_to_golden.py" "$REPO/oracle/_ref/dump" "$REPO/tests/golden/ref_hotpath.npz"
echo "wrote tests/golden/ref_hotpath.npz -- run: python -m pytest tests/test_oracle_ref_pin.py -q"
