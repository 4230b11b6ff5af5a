// oc_march2_inst.cu — instantiations of the two-columns-per-thread marching kernel, one object per mode.
#include "oc_march2.cuh"

#if OC_INST_EXACT
typedef MathExact OcInstMath;
extern "C" const void* oc_march2_fn_exact(int WC)
#else
typedef MathFast OcInstMath;
extern "C" const void* oc_march2_fn_fast(int WC)
#endif
{
    switch (WC) {
    case 64:  return (const void*)&oc_k_march2<OcInstMath, 64>;
    case 128: return (const void*)&oc_k_march2<OcInstMath, 128>;
    default:  return nullptr;
    }
}
