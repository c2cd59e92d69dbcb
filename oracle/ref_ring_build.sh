#!/bin/bash
# Builds the UNMODIFIED reference sources of its ring / proclog / affinity
# runtime, where they lie under /root/reference/src, into
# oracle/_ref/libbifrost_ref_ring.so -- CPU only (g++, no CUDA), so the
# differential tests of bifrost_b200/csrc/ring.cpp (tests/test_ring.py) run
# without a GPU.  TEST INFRASTRUCTURE: nothing under bifrost_b200/ links or
# loads it.  The reference's autotools build is NOT run; the only generated
# header these files need (bifrost/config.h) is hand-written below.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/src" ]; then
	echo "ref_ring_build: $REF not present; keeping whatever is in $OUT" >&2
	exit 0
fi
mkdir -p $OUT/include_cpu/bifrost $OUT/obj_ring
cat > $OUT/include_cpu/bifrost/config.h <<'EOF'
/* Hand-written stand-in for the autoconf-generated bifrost/config.h (CPU-only build) */
#ifndef BF_CONFIG_H_INCLUDE_GUARD_
#define BF_CONFIG_H_INCLUDE_GUARD_
#define BF_ALIGNMENT 4096
#define BF_CUDA_ENABLED 0
#define BF_GPU_MANAGEDMEM 0
#define BF_GPU_EXP_PINNED_ALLOC 0
#define BF_FLOAT128_ENABLED 0
#define BF_OPENMP_ENABLED 0
#define BF_HWLOC_ENABLED 0
#define BF_VMA_ENABLED 0
#define BF_VERBS_ENABLED 0
#define BF_RDMA_ENABLED 0
#define BF_DEBUG_ENABLED 0
#define BF_TRACE_ENABLED 0
#define BF_CUDA_DEBUG_ENABLED 0
#define BF_PROCLOG_DIR "/dev/shm/bifrost_ref_oracle"
#endif
EOF
# (not $CXX: this image points it at a g++ that links libstdc++ statically, and two
# copies of libstdc++ in one process share their GNU_UNIQUE locale guards -> crash)
CXX=${REF_CXX:-/usr/bin/g++}
OBJS=""
for s in ring ring_impl proclog fileutils affinity memory common cuda hw_locality; do
	o=$OUT/obj_ring/$s.o
	OBJS="$OBJS $o"
	$CXX -O2 -std=c++17 -fPIC -w -I$OUT/include_cpu -I$REF/src -c $REF/src/$s.cpp -o $o
done
$CXX -shared -o $OUT/libbifrost_ref_ring.so $OBJS -lpthread
echo "ref_ring_build: wrote $OUT/libbifrost_ref_ring.so"
