#!/bin/bash
python tools/rs_block.py
python tools/kernel_probe.py row_select 2>&1 | grep row_select | cut -c1-160
timeout 600 python -m pytest tests -x -q -m gpu -k "row" 2>&1 | tail -3
