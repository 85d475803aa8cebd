// Instantiations of the hot kernel (fast_kernel.cuh), one translation unit per SB_PART so they compile in parallel:
//   SB_PART = hs_index * 8 + cost * 4 + radius subset;  hs_index 0 -> 1 strip per warp, 1 -> 2 strips per warp;
//   cost 0 = SSD, 1 = NCC;  radius subsets {0,1,2,3}, {4}, {5}, {6,7}.
//   SB_PART = 16, 17, 18: the fused pair kernels (SSD, both maps of a pair from one cost volume), radius subsets
//   {0,1,2,3}, {4}, {5}.
#include "fast_kernel.cuh"

#ifndef SB_PART
#error "compile with -DSB_PART=0..18"
#endif

namespace sb {

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)

#if SB_PART >= 16
#if SB_PART == 16
fast_kernel_fn fast_pick_fused_a(int R) {
    switch (R) {
    case 0: return fast_cost_kernel<0, FK_FUSED, FWARPS, STEREO_COST_SSD, 1, true>;
    case 1: return fast_cost_kernel<1, FK_FUSED, FWARPS, STEREO_COST_SSD, 1, true>;
    case 2: return fast_cost_kernel<2, FK_FUSED, FWARPS, STEREO_COST_SSD, 1, true>;
    case 3: return fast_cost_kernel<3, FK_FUSED, FWARPS, STEREO_COST_SSD, 1, true>;
    }
    return nullptr;
}
#elif SB_PART == 17
fast_kernel_fn fast_pick_fused_b(int R) { return R == 4 ? fast_cost_kernel<4, FK_FUSED, FWARPS, STEREO_COST_SSD, 1, true> : nullptr; }
#else
fast_kernel_fn fast_pick_fused_c(int R) { return R == 5 ? fast_cost_kernel<5, FK_FUSED, FWARPS, STEREO_COST_SSD, 1, true> : nullptr; }
#endif
#else
// `key` = strips per warp | cost << 8
fast_kernel_fn SB_CAT(fast_pick_part, SB_PART)(int R, int key) {
    constexpr int HS = (SB_PART / 8) ? 2 : 1;
    constexpr int COST = ((SB_PART / 4) % 2) ? STEREO_COST_NCORR : STEREO_COST_SSD;
    constexpr int SUB = SB_PART % 4;
    if (key != (HS | (COST << 8))) return nullptr;
    if constexpr (SUB == 1) {
        if (R == 4) return fast_cost_kernel<4, FK_DEFAULT, FWARPS, COST, HS>;
    } else if constexpr (SUB == 2) {
        if (R == 5) return fast_cost_kernel<5, FK_DEFAULT, FWARPS, COST, HS>;
    } else if constexpr (SUB == 0) {
        switch (R) {
        case 0: return fast_cost_kernel<0, FK_DEFAULT, FWARPS, COST, HS>;
        case 1: return fast_cost_kernel<1, FK_DEFAULT, FWARPS, COST, HS>;
        case 2: return fast_cost_kernel<2, FK_DEFAULT, FWARPS, COST, HS>;
        case 3: return fast_cost_kernel<3, FK_DEFAULT, FWARPS, COST, HS>;
        }
    } else {
        if (R == 6) return fast_cost_kernel<6, FK_DEFAULT, FWARPS, COST, HS>;
        if (R == 7) return fast_cost_kernel<7, FK_DEFAULT, FWARPS, COST, HS>;
    }
    return nullptr;
}
#endif

} // namespace sb
