#!/bin/bash
for bx in 8 4 2 1; do echo "PR_REFIT_BATCH_BX=$bx"; PR_REFIT_BATCH_BX=$bx timeout 200 python ubench/batch_rate.py 2>&1 | tail -1; done
