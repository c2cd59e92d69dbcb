#!/bin/bash
# round-2 GPU check X (1 GPU, ~40 s): the executor's GPU tests after the last host-side edits
cd "$(dirname "$0")/../.."
timeout -s KILL 70 python -m pytest tests/test_blocks_gpu.py tests/test_io_formats.py tests/test_ring.py tests/test_cabi.py -m gpu -x -q 2>&1 | tail -4
