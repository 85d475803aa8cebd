// Instantiations of the hot kernel (fast_kernel.cuh), one translation unit per SB_PART so they compile in parallel:
//   SB_PART = hs_index * 8 + cost * 4 + radius subset;  hs_index 0 -> 1 strip per warp, 1 -> 2 strips per warp;
//   cost 0 = SSD, 1 = NCC;  radius subsets {0,1,2,3}, {4}, {5}, {6,7}.
//   SB_PART = 16..21: the fused pair kernels (SSD, both maps of a pair from one cost volume), radius subsets
//   {0,1,2,3}, {4}, {5}; 16-18 one strip per warp, 19-21 two strips per warp.
#include "fast_kernel.cuh"

#ifndef SB_PART
#error "compile with -DSB_PART=0..21"
#endif

namespace sb {

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)

#if SB_PART >= 16
#define SB_FUSED_KERNEL(R_, HS_) fast_cost_kernel<R_, (HS_ == 2 ? FK_FUSED2 : FK_FUSED), FWARPS, STEREO_COST_SSD, HS_, true>
#if SB_PART == 16 || SB_PART == 19
#if SB_PART == 16
fast_kernel_fn fast_pick_fused_a(int R) {
    constexpr int HS = 1;
#else
fast_kernel_fn fast_pick_fused2_a(int R) {
    constexpr int HS = 2;
#endif
    switch (R) {
    case 0: return SB_FUSED_KERNEL(0, HS);
    case 1: return SB_FUSED_KERNEL(1, HS);
    case 2: return SB_FUSED_KERNEL(2, HS);
    case 3: return SB_FUSED_KERNEL(3, HS);
    }
    return nullptr;
}
#elif SB_PART == 17
fast_kernel_fn fast_pick_fused_b(int R) { return R == 4 ? SB_FUSED_KERNEL(4, 1) : nullptr; }
#elif SB_PART == 18
fast_kernel_fn fast_pick_fused_c(int R) { return R == 5 ? SB_FUSED_KERNEL(5, 1) : nullptr; }
#elif SB_PART == 20
fast_kernel_fn fast_pick_fused2_b(int R) { return R == 4 ? SB_FUSED_KERNEL(4, 2) : nullptr; }
#else
fast_kernel_fn fast_pick_fused2_c(int R) { return R == 5 ? SB_FUSED_KERNEL(5, 2) : nullptr; }
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
