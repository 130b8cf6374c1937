#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "prepare_hinv or first_order or zeroth_order_loop" > gpurun_out/pytest_r2e.log 2>&1
tail -40 gpurun_out/pytest_r2e.log
