#!/bin/bash
# compute-sanitizer memcheck over a subset of the GPU tests (out-of-bounds shared / global accesses of the hot kernels)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -x --tb=line -k "${1:-fused_pair_any_range or float_fused or ncc_fused_pairs_vs or ssd_float_kernels_vs or ncc_float or refgpu_kernels or host_packing or ssd_fast_u8 or ncc_fast_u8}" > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize.log | sort | uniq -c | tail -15
