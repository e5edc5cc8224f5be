#!/bin/bash
# round 2, GPU call L: the two test files edited after the last full-suite run.
timeout 120 python -m pytest tests/test_gpu_reference_own_checks.py tests/test_gpu_next_rows.py -m gpu -q -x 2>&1 | tail -3
