// Instantiations of the hot kernel (fast_kernel.cuh), one translation unit per SB_PART so they compile in parallel:
//   SB_PART = 0..15: hs_index * 8 + cost * 4 + radius subset;  hs_index 0 -> 1 strip per warp, 1 -> 2 strips per warp;
//   cost 0 = SSD, 1 = NCC;  radius subsets {0,1,2,3}, {4}, {5}, {6,7}.
//   SB_PART = 16..25: the fused pair kernels (SSD, both maps of a pair from one cost volume):
//   16..20 one strip per warp, radius subsets {0,1,2,3}, {4}, {5}, {6}, {7};  21..25 two strips per warp, same subsets.
//   SB_PART = 26..37: the float-operand kernels (OPF): kind * 4 + radius subset {0,1,2,3}, {4}, {5}, {6,7};
//   kind 0 = SSD, 1 = SSD fused pair, 2 = NCC.
//   SB_PART = 38..47: the fused NCC pair kernels (packed operands), laid out like 16..25.
#include "fast_kernel.cuh"

#ifndef SB_PART
#error "compile with -DSB_PART=0..47"
#endif

namespace sb {

#define SB_CAT2(a, b) a##b
#define SB_CAT(a, b) SB_CAT2(a, b)

#if SB_PART >= 38
// fused NCC pair kernels: 38..42 one strip per warp, radius subsets {0,1,2,3}, {4}, {5}, {6}, {7}; 43..47 two strips per warp
#define SB_FUSED_NCC_KERNEL(R_, HS_) fast_cost_kernel<R_, FK_FUSED_NCC, FWARPS, STEREO_COST_NCORR, HS_, true, true>
fast_kernel_fn SB_CAT(fast_pick_fused_part, SB_PART)(int R, int hs, int gen) {
    constexpr int HS = (SB_PART - 38) / 5 + 1;
    constexpr int SUB = (SB_PART - 38) % 5;
    (void)gen;
    if (hs != HS) return nullptr;
    if constexpr (SUB == 0) {
        switch (R) {
        case 0: return SB_FUSED_NCC_KERNEL(0, HS);
        case 1: return SB_FUSED_NCC_KERNEL(1, HS);
        case 2: return SB_FUSED_NCC_KERNEL(2, HS);
        case 3: return SB_FUSED_NCC_KERNEL(3, HS);
        }
    } else {
        constexpr int RR = SUB + 3;       // 4, 5, 6, 7
        if (R == RR) return SB_FUSED_NCC_KERNEL(RR, HS);
    }
    return nullptr;
}
#elif SB_PART >= 26
// float-operand kernels (general float32 images): kind 0 = SSD, 1 = SSD fused pair, 2 = NCC
#define SB_OPF_KERNEL(R_) fast_cost_kernel<R_, fast_kf(R_), FWARPS, (KIND == OPF_NCC ? STEREO_COST_NCORR : STEREO_COST_SSD), 1, KIND == OPF_SSD_FUSED, true, true>
fast_kernel_fn SB_CAT(fast_pick_opf_part, SB_PART)(int R, int kind) {
    constexpr int KIND = (SB_PART - 26) / 4;
    constexpr int SUB = (SB_PART - 26) % 4;
    if (kind != KIND) return nullptr;
    if constexpr (SUB == 0) {
        switch (R) {
        case 0: return SB_OPF_KERNEL(0);
        case 1: return SB_OPF_KERNEL(1);
        case 2: return SB_OPF_KERNEL(2);
        case 3: return SB_OPF_KERNEL(3);
        }
    } else if constexpr (SUB == 1) {
        if (R == 4) return SB_OPF_KERNEL(4);
    } else if constexpr (SUB == 2) {
        if (R == 5) return SB_OPF_KERNEL(5);
    } else {
        if (R == 6) return SB_OPF_KERNEL(6);
        if (R == 7) return SB_OPF_KERNEL(7);
    }
    return nullptr;
}
#elif SB_PART >= 16
// gen = 0: the MODE-1-only kernel (R <= 5); gen = 1: the general one
#define SB_FUSED_KERNEL(R_, HS_, G_) fast_cost_kernel<R_, fast_k(R_, true, HS_), FWARPS, STEREO_COST_SSD, HS_, true, G_>
#define SB_FUSED_BOTH(R_, HS_) (gen ? SB_FUSED_KERNEL(R_, HS_, true) : SB_FUSED_KERNEL(R_, HS_, false))
fast_kernel_fn SB_CAT(fast_pick_fused_part, SB_PART)(int R, int hs, int gen) {
    constexpr int HS = (SB_PART - 16) / 5 + 1;
    constexpr int SUB = (SB_PART - 16) % 5;
    if (hs != HS) return nullptr;
    if constexpr (SUB == 0) {
        switch (R) {
        case 0: return SB_FUSED_BOTH(0, HS);
        case 1: return SB_FUSED_BOTH(1, HS);
        case 2: return SB_FUSED_BOTH(2, HS);
        case 3: return SB_FUSED_BOTH(3, HS);
        }
    } else if constexpr (SUB <= 2) {
        constexpr int RR = SUB + 3;       // 4, 5
        if (R == RR) return SB_FUSED_BOTH(RR, HS);
    } else {
        constexpr int RR = SUB + 3;       // 6, 7: MODE 1 does not exist without the other flavours (explicit border selects)
        if (R == RR && gen) return SB_FUSED_KERNEL(RR, HS, true);
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
