#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_last.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_last.log
tail -3 gpurun_out/pytest_gpu_last.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_last.log
