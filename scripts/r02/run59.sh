#!/bin/bash
timeout 600 python -m pytest tests/test_iaf_tc_gpu.py -m gpu -q -x --timeout 300 -k "pair" 2>&1 | tail -6
