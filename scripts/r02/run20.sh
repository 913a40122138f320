#!/bin/bash
timeout 600 python scripts/r02/chunk_diag.py 2>&1 | tail -12
