#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest.log; tail -4 gpurun_out/r02q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
