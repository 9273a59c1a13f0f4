#!/bin/bash
# round 2, call C (N GPUs): multi-device tests, bench under torchrun, optional H2D topology experiment
N=${1:-2}
out=gpurun_out/r2c_n$N; mkdir -p $out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $out/smi.txt 2>&1; nproc >> $out/smi.txt; free -g >> $out/smi.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "sharded" > $out/tests_sharded.log 2>&1; echo "rc=$?" >> $out/tests_sharded.log
tail -4 $out/tests_sharded.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err ) 2> $out/bench_n$N.time
tail -c 1500 $out/bench_n$N.err; cat $out/bench_n$N.time
python tools/benchline.py "N=$N" < $out/bench_n$N.json
if [ "$2" = "topo" ]; then
  timeout 600 python tools/h2d_topology.py > $out/h2d_topology.txt 2>&1
  tail -20 $out/h2d_topology.txt
fi
