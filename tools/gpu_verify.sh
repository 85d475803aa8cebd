#!/bin/bash
# last check of the tree as committed: GPU tests, smoke, the default bench line, PCIe bandwidth of the box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1i_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1i_smoke.log
timeout 60 python tools/pcie_bw.py > gpurun_out/r1i_pcie_bw.json 2>&1; cat gpurun_out/r1i_pcie_bw.json
timeout 300 python bench.py > gpurun_out/r1i_bench_n1_4k_ssd.json 2> gpurun_out/r1i_err.log; echo "bench rc=$?"
