#!/bin/bash
O=gpurun_out/r1e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or host_buffer or mixed_corpus or gate or empty_and_ragged or save_batch" 2>&1 | tail -5 | tee $O/pytest_pipe.log
python bench.py --no-cpu-baseline > $O/bench_c2_pipe.json 2> $O/bench_c2_pipe.err; tail -3 $O/bench_c2_pipe.err
python tools/benchline.py c2pipe < $O/bench_c2_pipe.json
python -c "
import json;d=json.load(open('$O/bench_c2_pipe.json'));print(d['e2e'])"
