#!/bin/bash
mkdir -p gpurun_out
echo "=== policy replay"; timeout 900 python -m pytest tests/test_policy_replay.py -m gpu -q -s > gpurun_out/pytest_policy.log 2>&1; tail -12 gpurun_out/pytest_policy.log | cut -c1-300
