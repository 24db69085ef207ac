#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_cpp_host.py -q -m gpu --tb=short 2>&1 | tail -15
