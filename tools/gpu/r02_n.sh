#!/bin/bash
# round-2 GPU check N (1 GPU): sharded full-band FDMT emulated on one device, new defaults
timeout -s KILL 900 python -m pytest tests/test_fdmt_sharded.py -x -q -m gpu 2>&1 | tail -15
timeout -s KILL 300 python tools/fdmt_time.py "" 2>&1 | tail -2
