#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.  Builds the verbatim reference physics into
# oracle/_ref/libocref.so from the sources WHERE THEY LIE under /root/reference.
#
# The reference (OpenCloth_Verlet/OpenCloth_Verlet/main.cpp) cannot be built with its own
# build system (Visual Studio project, GLUT/GLEW/Win32).  Its physics, however, is self
# contained: this script cuts the physics line ranges out of main.cpp with sed into
# oracle/_ref/slices/ (git-ignored) and compiles oracle/ref_shim.cpp, which #includes them,
# against the reference's vendored GLM 0.9.0.0.
#
#   -ffp-contract=off        no FMA contraction (x86-64 baseline has none anyway)
#   -Dglm_core_func_integer  pre-defines the include guard of glm/core/func_integer.hpp; its
#                            .inl (lines 149-150, 211-212) does not compile with g++ 13 and
#                            nothing on the path uses integer functions
#   -fpermissive -w          2010-era GLM under a 2024 compiler
#
# Outputs ONLY into oracle/_ref/.  No-op (exit 0) if /root/reference is absent (GPU box): the
# prebuilt oracle/_ref/libocref.so travels with the snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${OC_REFERENCE_ROOT:-/root/reference}"
SRC="$REF/OpenCloth_Verlet/OpenCloth_Verlet/main.cpp"
OUT="$HERE/_ref"
if [ ! -f "$SRC" ]; then
  echo "build_ref: $SRC not present; keeping prebuilt $OUT/libocref.so (if any)"
  exit 0
fi
mkdir -p "$OUT/slices"
cut_slice() { sed -n "$1,$2p" "$SRC" | tr -d '\r' > "$OUT/slices/$3"; }
cut_slice  67  72 spring_struct.inc       # struct Spring
cut_slice  90  93 constants.inc           # STRUCTURAL/SHEAR/BEND ids, spring_count
cut_slice  97 104 params.inc              # DEFAULT_DAMPING, Ks*/Kd*, gravity, mass, timeStep
cut_slice 123 132 ellipsoid_globals.inc   # ellipsoid matrices, center, radius, StepPhysics decl
cut_slice 134 144 add_spring.inc          # AddSpring
cut_slice 253 260 init_positions.inc      # InitGL: positions
cut_slice 286 327 init_springs.inc        # InitGL: springs + ellipsoid matrices
cut_slice 428 484 physics.inc             # IntegrateVerlet, GetVerletVelocity, ComputeForces
cut_slice 509 533 collision.inc           # EllipsoidCollision
cut_slice 557 562 step.inc                # StepPhysics
g++ -std=gnu++11 -O2 -ffp-contract=off -fpermissive -w -Dglm_core_func_integer \
    -shared -fPIC -I"$REF/dep/glm" -I"$HERE" "$HERE/ref_shim.cpp" -o "$OUT/libocref.so"
echo "build_ref: built $OUT/libocref.so"
