// Instantiations of the hot kernel (fast_kernel.cuh), one translation unit per SB_PART so they compile in parallel:
//   SB_PART = 0..15: hs_index * 8 + cost * 4 + radius subset;  hs_index 0 -> 1 strip per warp, 1 -> 2 strips per warp;
//   cost 0 = SSD, 1 = NCC;  radius subsets {0,1,2,3}, {4}, {5}, {6,7}.
//   SB_PART = 16..25: the fused pair kernels (SSD, both maps of a pair from one cost volume):
//   16..20 one strip per warp, radius subsets {0,1,2,3}, {4}, {5}, {6}, {7};  21..25 two strips per warp, same subsets.
#include "fast_kernel.cuh"

#ifndef SB_PART
#error "compile with -DSB_PART=0..25"
#endif

namespace sb {

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)

#if SB_PART >= 16
#define SB_FUSED_KERNEL(R_, HS_) fast_cost_kernel<R_, fast_k(R_, true, HS_), FWARPS, STEREO_COST_SSD, HS_, true>
fast_kernel_fn SB_CAT(fast_pick_fused_part, SB_PART)(int R, int hs) {
    constexpr int HS = (SB_PART - 16) / 5 + 1;
    constexpr int SUB = (SB_PART - 16) % 5;
    if (hs != HS) return nullptr;
    if constexpr (SUB == 0) {
        switch (R) {
        case 0: return SB_FUSED_KERNEL(0, HS);
        case 1: return SB_FUSED_KERNEL(1, HS);
        case 2: return SB_FUSED_KERNEL(2, HS);
        case 3: return SB_FUSED_KERNEL(3, HS);
        }
    } else {
        constexpr int RR = SUB + 3;       // 4, 5, 6, 7
        if (R == RR) return SB_FUSED_KERNEL(RR, HS);
    }
    return nullptr;
}
#else
// `key` = strips per warp | cost << 8
fast_kernel_fn SB_CAT(fast_pick_part, SB_PART)(int R, int key) {
    constexpr int HS = (SB_PART / 8) ? 2 : 1;
    constexpr int COST = ((SB_PART / 4) % 2) ? STEREO_COST_NCORR : STEREO_COST_SSD;
    constexpr int SUB = SB_PART % 4;
    if (key != (HS | (COST << 8))) return nullptr;
#define SB_KERNEL(R_) fast_cost_kernel<R_, fast_k(R_, false, HS), FWARPS, COST, HS>
    if constexpr (SUB == 1) {
        if (R == 4) return SB_KERNEL(4);
    } else if constexpr (SUB == 2) {
        if (R == 5) return SB_KERNEL(5);
    } else if constexpr (SUB == 0) {
        switch (R) {
        case 0: return SB_KERNEL(0);
        case 1: return SB_KERNEL(1);
        case 2: return SB_KERNEL(2);
        case 3: return SB_KERNEL(3);
        }
    } else {
        if (R == 6) return SB_KERNEL(6);
        if (R == 7) return SB_KERNEL(7);
    }
    return nullptr;
}
#endif

} // namespace sb
