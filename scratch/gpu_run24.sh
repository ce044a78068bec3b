#!/bin/bash
timeout 300 python scratch/ssd_bwd_dbg.py 2>&1 | grep -i "per-CTA\|pieces\|warp 0" | head -12
