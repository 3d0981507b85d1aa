#!/bin/bash
timeout 400 python -m pytest tests/test_distributed.py -m gpu -q -x 2>&1 | tail -12
