#!/bin/bash
# Build the UNMODIFIED reference CPU path (LLNL/ddcMD, /root/reference/src) as the
# parity oracle.  TEST INFRASTRUCTURE ONLY.
#
# Sources are compiled where they lie; outputs go only into oracle/_ref/ (git-ignored,
# but shipped to the GPU box with the snapshot).  No reference source is copied.
# The reference's own build system (cmake + MPI + FFTW + OpenMP) is not run: the CPU
# path needs none of those except <mpi.h>, which oracle/mpi_shim supplies (1 rank).
#
# Flags follow CMakeLists.txt:32 of the reference (-DWITH_MPI -DWITH_PIO -D_GNU_SOURCE
# -DSYSTEM_LINUX), USE_GPU undefined => GPUCODE(x) is empty (src/HAVEGPU.h:10-12).
# -ffp-contract=off and no -march: the oracle must not fuse multiply-adds, so its
# pair-list membership test is the plain IEEE sequence the CUDA path reproduces.
#
# Produces:
#   oracle/_ref/ddcMD_ref   stock binary (main() from src/ddcMD.c) - the CPU baseline
#   oracle/_ref/ref_dump    oracle/ref_dump.c linked against the same objects
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF" ]; then
    echo "build_ref: $REF not present; keeping prebuilt $OUT" >&2
    exit 0
fi
if [ -x "$OUT/ddcMD_ref" ] && [ -x "$OUT/ref_dump" ] && [ "$OUT/ref_dump" -nt "$HERE/ref_dump.c" ] \
   && [ "$OUT/ref_dump" -nt "$HERE/mpi_shim/mpi_stub.c" ] && [ -z "${FORCE:-}" ]; then
    echo "build_ref: up to date"
    exit 0
fi
mkdir -p "$OBJ"
CFLAGS="-std=gnu99 -O2 -w -ffp-contract=off -DWITH_MPI -DWITH_PIO -D_GNU_SOURCE -DSYSTEM_LINUX -I$HERE/mpi_shim -I$REF"

cat > "$OBJ/Makefile" <<EOF
REF=$REF
CFLAGS=$CFLAGS
SRCS=\$(wildcard \$(REF)/*.c)
OBJS=\$(patsubst \$(REF)/%.c,%.o,\$(SRCS))
all: \$(OBJS) binProcess.o ddcMD_testexe.o mpi_stub.o ref_dump.o
%.o: \$(REF)/%.c
	gcc \$(CFLAGS) -c \$< -o \$@
binProcess.o: \$(REF)/binProcess.cpp
	g++ -O2 -w -ffp-contract=off -DWITH_MPI -DWITH_PIO -D_GNU_SOURCE -DSYSTEM_LINUX -I$HERE/mpi_shim -I\$(REF) -c \$< -o \$@
ddcMD_testexe.o: \$(REF)/ddcMD.c
	gcc \$(CFLAGS) -DTESTEXE=1 -c \$< -o \$@
mpi_stub.o: $HERE/mpi_shim/mpi_stub.c
	gcc -O2 -I$HERE/mpi_shim -c \$< -o \$@
ref_dump.o: $HERE/ref_dump.c
	gcc \$(CFLAGS) -c \$< -o \$@
EOF
make -C "$OBJ" -j"$JOBS" -s all

cd "$OBJ"
LIBOBJS=$(ls *.o | grep -v -e '^ddcMD.o$' -e '^ddcMD_testexe.o$' -e '^ref_dump.o$')
g++ -o "$OUT/ddcMD_ref" ddcMD.o $LIBOBJS -lm -lpthread
g++ -o "$OUT/ref_dump" ref_dump.o ddcMD_testexe.o $LIBOBJS -lm -lpthread
echo "build_ref: built $OUT/ddcMD_ref and $OUT/ref_dump"
