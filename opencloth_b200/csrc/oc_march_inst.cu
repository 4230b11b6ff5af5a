// oc_march_inst.cu — explicit instantiations of the marching kernel.  Compiled once per
// (OC_INST_TW, OC_INST_EXACT) pair so that the variants build in parallel:
//   TW = 32, 64 : S = 1, 2, 4, 8      TW = 128 : S = 1, 2, 4
// Resident CTAs per SM the 128-thread kernels are register-capped for: the exact kernels need ~160
// registers to hold three spring pairs without spilling (3 CTAs), the fast ones fit 128 (4 CTAs).
// Measured on B200 at 2048^2: exact 3 CTAs 32.5 vs 4 CTAs 30.5 G updates/s; 5 CTAs (96 regs, spills) is
// slower in both modes.
#if OC_INST_EXACT
#define OC_CTAS128 3
#else
#define OC_CTAS128 4
#endif
#include "oc_march.cuh"

#ifndef OC_INST_TW
#error "compile with -DOC_INST_TW=32|64|128 -DOC_INST_EXACT=0|1"
#endif

#if OC_INST_EXACT
typedef MathExact OcInstMath;
#define OC_INST_NAME2(tw) oc_march_fn_exact_##tw
#else
typedef MathFast OcInstMath;
#define OC_INST_NAME2(tw) oc_march_fn_fast_##tw
#endif
#define OC_INST_NAME1(tw) OC_INST_NAME2(tw)
#define OC_INST_NAME OC_INST_NAME1(OC_INST_TW)

// returns the kernel for S stages, or nullptr when that stage count is not built for this width
extern "C" const void* OC_INST_NAME(int S)
{
    switch (S) {
    case 1: return (const void*)&oc_k_march<OcInstMath, 1, OC_INST_TW>;
    case 2: return (const void*)&oc_k_march<OcInstMath, 2, OC_INST_TW>;
    case 4: return (const void*)&oc_k_march<OcInstMath, 4, OC_INST_TW>;
#if OC_INST_TW <= 64
    case 8: return (const void*)&oc_k_march<OcInstMath, 8, OC_INST_TW>;
#endif
    default: return nullptr;
    }
}
