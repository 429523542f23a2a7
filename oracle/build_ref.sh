#!/bin/sh
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
# Compiles the UNMODIFIED reference sources where they lie under $REF
# (default /root/reference) into oracle/_ref/<flavour>/ with plain g++/gcc
# calls -- the reference's own build system is not run and no reference
# source is copied into this repository.
#
#   flavour "fast"   = the flags the reference ships (setup_gcc.sh:9)   -> Oracle-F
#   flavour "strict" = no fast-math, no FMA contraction                 -> Oracle-S
#
# Outputs per flavour: libminrms_ref.a (all lib/* objects), minrms (the CLI), and the
# drivers ref_dump / ref_bench / ref_plot (oracle/*.cc linked against the reference objects).
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$REF/lib/dyn" ] || { echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref" >&2; exit 0; }

LIBDIRS="global_utils parse_utils vect_3d eigen pdb++ fast_rot_metric biopolymer dyn superimpose or_gen pair_alignment"
INC=""
for d in $LIBDIRS; do INC="$INC -I$REF/lib/$d"; done
INC="$INC -I$REF/bin/minrms"

build_flavour() {
  flav=$1; shift
  FLAGS="$*"
  O=$OUT/$flav
  mkdir -p $O/obj
  objs=""
  for d in $LIBDIRS; do
    for src in $REF/lib/$d/*.cc $REF/lib/$d/*.c; do
      [ -f "$src" ] || continue
      base=$(basename $src)
      case $base in
        demoNW.cc|binary_file_demo.cc|pythag.c) continue;;
      esac
      obj=$O/obj/${d}_${base%.*}.o
      objs="$objs $obj"
      if [ ! -f $obj ] || [ $src -nt $obj ]; then
        case $src in
          *.c)  gcc $FLAGS -w -c $INC $src -o $obj & ;;
          *)    g++ -std=c++11 $FLAGS -w -c $INC $src -o $obj & ;;
        esac
      fi
    done
    wait
  done
  rm -f $O/libminrms_ref.a
  ar rcs $O/libminrms_ref.a $objs
  binobjs=""
  for src in $REF/bin/minrms/*.cc; do
    base=$(basename $src .cc)
    obj=$O/obj/bin_${base}.o
    [ "$base" = main ] || binobjs="$binobjs $obj"
    if [ ! -f $obj ] || [ $src -nt $obj ]; then
      g++ -std=c++11 $FLAGS -w -c $INC $src -o $obj &
    fi
  done
  wait
  rm -f $O/libminrms_bin.a
  ar rcs $O/libminrms_bin.a $binobjs
  g++ -o $O/minrms $O/obj/bin_main.o $O/libminrms_bin.a $O/libminrms_ref.a
  for drv in ref_dump ref_bench ref_plot; do
    if [ -f $HERE/$drv.cc ]; then
      g++ -std=c++11 $FLAGS -w $INC $HERE/$drv.cc $O/libminrms_bin.a $O/libminrms_ref.a -o $O/$drv
    fi
  done
  # the drop-in binary: reference CLI + SearchOrCuda (minrms_b200/host) over libminrms_b200.so
  HOST=$HERE/../minrms_b200/host
  LIBDIR=$HERE/../minrms_b200
  if [ -f $LIBDIR/libminrms_b200.so ]; then
    g++ -std=c++11 $FLAGS -w $INC -I$HERE/../include $HOST/main_cuda.cc $HOST/search_or_cuda.cc \
        $O/libminrms_bin.a $O/libminrms_ref.a -L$LIBDIR -lminrms_b200 -lpthread \
        -Wl,-rpath,'$ORIGIN/../../../minrms_b200' -o $O/minrms_cuda
    # all-vs-all driver over the same objects (one process per pair, see batch_cuda.cc)
    g++ -std=c++11 $FLAGS -w $INC -I$HERE/../include $HOST/batch_cuda.cc $HOST/search_or_cuda.cc \
        $O/libminrms_bin.a $O/libminrms_ref.a -L$LIBDIR -lminrms_b200 -lpthread \
        -Wl,-rpath,'$ORIGIN/../../../minrms_b200' -o $O/minrms_batch_cuda
  fi
  echo "built $O"
}

build_flavour fast   -O3 -DNDEBUG -ffast-math
build_flavour strict -O2 -DNDEBUG -ffp-contract=off
