#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/all_v6.log
