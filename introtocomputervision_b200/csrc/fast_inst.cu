// Instantiations of the hot kernel (fast_kernel.cuh), one translation unit per SB_PART so they compile in parallel:
//   parts 0-3: SSD, parts 4-7: NCC;  radius subsets {0,1,2,3}, {4}, {5}, {6,7};  K=16 variants for R = 4, 5.
#include "fast_kernel.cuh"

#ifndef SB_PART
#error "compile with -DSB_PART=0..7"
#endif

namespace sb {

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)

fast_kernel_fn SB_CAT(fast_pick_part, SB_PART)(int R, int K) {
    constexpr int COST = (SB_PART < 4) ? STEREO_COST_SSD : STEREO_COST_NCORR;
    constexpr int SUB = SB_PART % 4;
    if constexpr (SUB == 1) {
        if (K == 16 && R == 4) return fast_cost_kernel<4, 16, 12, COST>;
        if (K == 24 && R == 4) return fast_cost_kernel<4, 24, 8, COST>;
    } else if constexpr (SUB == 2) {
        if (K == 16 && R == 5) return fast_cost_kernel<5, 16, 12, COST>;
        if (K == 24 && R == 5) return fast_cost_kernel<5, 24, 8, COST>;
    } else if constexpr (SUB == 0) {
        if (K == 24) switch (R) {
        case 0: return fast_cost_kernel<0, 24, 8, COST>;
        case 1: return fast_cost_kernel<1, 24, 8, COST>;
        case 2: return fast_cost_kernel<2, 24, 8, COST>;
        case 3: return fast_cost_kernel<3, 24, 8, COST>;
        }
    } else {
        if (K == 24 && R == 6) return fast_cost_kernel<6, 24, 8, COST>;
        if (K == 24 && R == 7) return fast_cost_kernel<7, 24, 8, COST>;
    }
    return nullptr;
}

} // namespace sb
