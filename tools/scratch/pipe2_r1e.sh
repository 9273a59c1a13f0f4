#!/bin/bash
O=gpurun_out/r1e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu_final.log
python tools/scratch/files_breakdown.py 2>&1 | tee $O/files_breakdown2.txt
