// oc_stream_inst.cu — instantiations of the streaming gather kernel, one object per mode.
// occ = resident CTAs per SM the registers are capped for (0: the default of the width).
#include "oc_stream.cuh"

#if OC_INST_EXACT
typedef MathExact OcInstMath;
extern "C" const void* oc_stream_fn_exact(int WC, int occ)
#else
typedef MathFast OcInstMath;
extern "C" const void* oc_stream_fn_fast(int WC, int occ)
#endif
{
    if (WC == 64) {
        switch (occ) {
        case 4: return (const void*)&oc_k_stream<OcInstMath, 64, 4>;
        case 0: case 6: return (const void*)&oc_k_stream<OcInstMath, 64, 6>;
        case 8: return (const void*)&oc_k_stream<OcInstMath, 64, 8>;
        default: return nullptr;
        }
    }
    if (WC == 128) {
        switch (occ) {
        case 2: return (const void*)&oc_k_stream<OcInstMath, 128, 2>;
        case 0: case 3: return (const void*)&oc_k_stream<OcInstMath, 128, 3>;
        case 4: return (const void*)&oc_k_stream<OcInstMath, 128, 4>;
        default: return nullptr;
        }
    }
    return nullptr;
}
