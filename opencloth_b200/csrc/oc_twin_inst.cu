// oc_twin_inst.cu — instantiations of the twin-tile marching kernel, one object per mode.
// occ = resident CTAs per SM the registers are capped for (0: the default, 2 x 128 / 4 x 64 threads at 255 registers).
#include "oc_twin.cuh"

#if OC_INST_EXACT
typedef MathExact OcInstMath;
extern "C" const void* oc_twin_fn_exact(int WC, int occ)
#else
typedef MathFast OcInstMath;
extern "C" const void* oc_twin_fn_fast(int WC, int occ)
#endif
{
    if (WC == 64) {
        switch (occ) {
        case 0: case 4: return (const void*)&oc_k_twin<OcInstMath, 64, 4>;
#ifdef OC_ALL_VARIANTS          // the register-capped builds measured in DESIGN.md 4.2 (slower; not part of the default build)
        case 5: return (const void*)&oc_k_twin<OcInstMath, 64, 5>;
        case 6: return (const void*)&oc_k_twin<OcInstMath, 64, 6>;
#endif
        default: return nullptr;
        }
    }
    if (WC == 128) {
        switch (occ) {
        case 0: case 2: return (const void*)&oc_k_twin<OcInstMath, 128, 2>;
#ifdef OC_ALL_VARIANTS
        case 3: return (const void*)&oc_k_twin<OcInstMath, 128, 3>;
#endif
        default: return nullptr;
        }
    }
    return nullptr;
}
