#!/usr/bin/env bash
# Round-2 GPU session R (2 GPUs): the whole GPU suite (single-GPU tests + slab tests at 2 ranks) on the final commit.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
tail -8 gpurun_out/r2r_pytest.log
