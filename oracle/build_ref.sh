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
cut_slice 486 508 provot.inc              # ApplyProvotDynamicInverse (disabled in V:561; ref_step_provot enables it)
cut_slice 557 562 step.inc                # StepPhysics
CXXFLAGS="-std=gnu++11 -O2 -ffp-contract=off -fpermissive -w -Dglm_core_func_integer -shared -fPIC"
g++ $CXXFLAGS -I"$REF/dep/glm" -I"$HERE" "$HERE/ref_shim.cpp" -o "$OUT/libocref.so"
echo "build_ref: built $OUT/libocref.so"

# ---- the sibling explicit integrators: same recipe, their own files ---------------------------
# OpenCloth_ExplicitEuler ("E:") and OpenCloth_SemiImplicit ("S:"): state X, V; ComputeForces on the stored
# velocities, IntegrateEuler / IntegrateSemiImplicit, EllipsoidCollision (zeroes V), ApplyProvotDynamicInverse (on V).
SRC="$REF/OpenCloth_ExplicitEuler/OpenCloth_ExplicitEuler/main.cpp"
mkdir -p "$OUT/slices_euler"
cut_slice() { sed -n "$1,$2p" "$SRC" | tr -d '\r' > "$OUT/slices_euler/$3"; }
cut_slice  64  64 timestep.inc            # timeStep
cut_slice  69  74 spring_struct.inc       # struct Spring
cut_slice  92  95 constants.inc           # spring type ids, spring_count
cut_slice  97 102 params.inc              # DEFAULT_DAMPING, Ks*/Kd*, gravity, mass
cut_slice 121 130 ellipsoid_globals.inc   # ellipsoid matrices, center, radius, StepPhysics decl
cut_slice 132 142 add_spring.inc          # AddSpring
cut_slice 262 270 init_state.inc          # InitGL: X, V
cut_slice 296 336 init_springs.inc        # InitGL: springs + ellipsoid matrices
cut_slice 434 466 forces.inc              # ComputeForces
cut_slice 469 482 integrate.inc           # IntegrateEuler
cut_slice 554 577 provot.inc              # ApplyProvotDynamicInverse
cut_slice 578 602 collision.inc           # EllipsoidCollision
g++ $CXXFLAGS -DOC_REF_EXPLICIT_EULER -I"$REF/dep/glm" -I"$OUT/slices_euler" "$HERE/ref_shim_euler.cpp" -o "$OUT/libocref_euler.so"
echo "build_ref: built $OUT/libocref_euler.so"

SRC="$REF/OpenCloth_SemiImplicit/OpenCloth_SemiImplicit/main.cpp"
mkdir -p "$OUT/slices_semi"
cut_slice() { sed -n "$1,$2p" "$SRC" | tr -d '\r' > "$OUT/slices_semi/$3"; }
cut_slice  45  45 timestep.inc
cut_slice  52  57 spring_struct.inc
cut_slice  73  78 constants.inc
cut_slice  80  85 params.inc
cut_slice  98 106 ellipsoid_globals.inc
cut_slice 108 118 add_spring.inc
cut_slice 228 237 init_state.inc
cut_slice 263 303 init_springs.inc
cut_slice 402 436 forces.inc              # ComputeForces
cut_slice 437 462 provot.inc              # ApplyProvotDynamicInverse
cut_slice 464 477 integrate.inc           # IntegrateSemiImplicit
cut_slice 478 502 collision.inc           # EllipsoidCollision
g++ $CXXFLAGS -DOC_REF_SEMI_IMPLICIT -I"$REF/dep/glm" -I"$OUT/slices_semi" "$HERE/ref_shim_euler.cpp" -o "$OUT/libocref_semi.so"
echo "build_ref: built $OUT/libocref_semi.so"

# ---- UpdateNormals of the lit demo: the render hand-off's normals ------------------------------
SRC="$REF/OpenCloth_ExplicitEuler_TextureMapped_Lit/OpenCloth_ExplicitEuler_TextureMapped_Lit/main.cpp"
mkdir -p "$OUT/slices_lit"
cut_slice() { sed -n "$1,$2p" "$SRC" | tr -d '\r' > "$OUT/slices_lit/$3"; }
cut_slice  89  90 vertex_struct.inc       # struct Vertex, vertices
cut_slice 312 328 init_indices.inc        # InitGL: triangle list
cut_slice 684 707 update_normals.inc      # UpdateNormals
g++ $CXXFLAGS -I"$REF/dep/glm" -I"$OUT/slices_lit" "$HERE/ref_shim_normals.cpp" -o "$OUT/libocref_normals.so"
echo "build_ref: built $OUT/libocref_normals.so"
