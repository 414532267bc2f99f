#!/bin/bash
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED" | tail -3; done
