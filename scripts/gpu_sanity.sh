#!/bin/bash
# last look at the final build: smoke() and the partitioned / two-CTA parity tests
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 150 python -m pytest tests/test_gpu_pipeline.py -q --timeout 100 -k "partitioned or two_scoring or batches" 2>&1 | tail -2
