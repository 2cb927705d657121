#!/bin/bash
# after the trait variants' event counters (ABI 10): whole GPU suite + smoke on the new build
T=gpurun_out/r02an
mkdir -p $T
python -m pytest tests -m gpu -q -x > $T/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $T/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $T/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $T/smoke.log
