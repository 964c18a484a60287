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
SHIM="$HERE/../integration"
if [ -x "$OUT/ddcMD_ref" ] && [ -x "$OUT/ref_dump" ] && [ "$OUT/ref_dump" -nt "$HERE/ref_dump.c" ] \
   && [ "$OUT/ref_dump" -nt "$HERE/mpi_shim/mpi_stub.c" ] && [ -x "$OUT/ddcMD_shim" ] && [ "$OUT/ddcMD_shim" -nt "$SHIM/ddcmd_shim.c" ] \
   && [ "$OUT/ddcMD_shim" -nt "$SHIM/ddcmd_shim_main.c" ] && [ "$OUT/ddcMD_shim" -nt "$HERE/../include/ddcmd_b200_host.h" ] \
   && { [ ! -e "$HERE/../tests/cpu_emu/libddcmd_b200_emu.so" ] || [ -x "$OUT/ddcMD_shim_emu" ]; } && [ -z "${FORCE:-}" ]; then
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
all: \$(OBJS) binProcess.o ddcMD_testexe.o mpi_stub.o ref_dump.o ddcmd_shim.o ddcmd_shim_main.o
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
ddcmd_shim.o: $SHIM/ddcmd_shim.c $HERE/../include/ddcmd_b200_host.h $HERE/../include/ddcmd_b200.h
	gcc \$(CFLAGS) -c \$< -o \$@
ddcmd_shim_main.o: $SHIM/ddcmd_shim_main.c
	gcc \$(CFLAGS) -c \$< -o \$@
EOF
make -C "$OBJ" -j"$JOBS" -s all

cd "$OBJ"
LIBOBJS=$(ls *.o | grep -v -e '^ddcMD.o$' -e '^ddcMD_testexe.o$' -e '^ref_dump.o$' -e '^ddcmd_shim.o$' -e '^ddcmd_shim_main.o$')
g++ -o "$OUT/ddcMD_ref" ddcMD.o $LIBOBJS -lm -lpthread
g++ -o "$OUT/ref_dump" ref_dump.o ddcMD_testexe.o $LIBOBJS -lm -lpthread
# ddcMD with the library plugged in at its plug-in seam (integration/ddcmd_shim.c): the reference's objects + the shim, linked
# against the product library (runs on the GPU box) and, when it has been built, against the CPU emulation of the kernels
# (tests/cpu_emu, test infrastructure) so the seam is exercised in the build container too.  rpaths are relative to the binary.
PROD="$HERE/../ddcmd_b200"
if [ -e "$PROD/libddcmd_b200.so" ]; then
    g++ -o "$OUT/ddcMD_shim" ddcmd_shim_main.o ddcmd_shim.o ddcMD_testexe.o $LIBOBJS -L"$PROD" -lddcmd_b200 \
        -Wl,-rpath,'$ORIGIN/../../ddcmd_b200' -Wl,--allow-shlib-undefined -lm -lpthread
fi
EMU="$HERE/../tests/cpu_emu"
if [ -e "$EMU/libddcmd_b200_emu.so" ]; then
    g++ -o "$OUT/ddcMD_shim_emu" ddcmd_shim_main.o ddcmd_shim.o ddcMD_testexe.o $LIBOBJS -L"$EMU" -lddcmd_b200_emu \
        -Wl,-rpath,'$ORIGIN/../../tests/cpu_emu' -lm -lpthread
fi
echo "build_ref: built $OUT/ddcMD_ref, $OUT/ref_dump and the ddcMD_shim binaries"
