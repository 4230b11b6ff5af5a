// oc_stream_inst.cu — instantiations of the streaming gather kernel, one object per mode.
// occ = resident CTAs per SM the registers are capped for (0: the default of the width).
#include "oc_stream.cuh"
#include "oc_stream2.cuh"

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
        case 0: case 6: return (const void*)&oc_k_stream<OcInstMath, 64, 6>;
#ifdef OC_ALL_VARIANTS          // the other register caps measured in DESIGN.md 4.2 (equal or slower; not part of the default build)
        case 4: return (const void*)&oc_k_stream<OcInstMath, 64, 4>;
        case 5: return (const void*)&oc_k_stream<OcInstMath, 64, 5>;
        case 8: return (const void*)&oc_k_stream<OcInstMath, 64, 8>;
#endif
        default: return nullptr;
        }
    }
    if (WC == 128) {
        switch (occ) {
        case 0: case 3: return (const void*)&oc_k_stream<OcInstMath, 128, 3>;
#ifdef OC_ALL_VARIANTS
        case 2: return (const void*)&oc_k_stream<OcInstMath, 128, 2>;
        case 4: return (const void*)&oc_k_stream<OcInstMath, 128, 4>;
#endif
        default: return nullptr;
        }
    }
    return nullptr;
}

// oc_k_stream2 (two columns per thread, 64 threads per 128-column window)
#if OC_INST_EXACT
extern "C" const void* oc_stream2_fn_exact(int WC, int occ)
#else
extern "C" const void* oc_stream2_fn_fast(int WC, int occ)
#endif
{
    if (WC != 128) return nullptr;
    switch (occ) {
    case 0: case 4: return (const void*)&oc_k_stream2<OcInstMath, 128, 4>;      // 228 registers; capped at 168 it spills and loses 10 %
#ifdef OC_ALL_VARIANTS
    case 5: return (const void*)&oc_k_stream2<OcInstMath, 128, 5>;
    case 6: return (const void*)&oc_k_stream2<OcInstMath, 128, 6>;
#endif
    default: return nullptr;
    }
}
