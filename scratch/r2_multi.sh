#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2mg; mkdir -p $o
nvidia-smi -L > $o/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $o/bench_n2.json 2> $o/bench_n2.err; echo "bench rc=$?"; tail -c 600 $o/bench_n2.json; tail -3 $o/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scratch/search_sharded.py 2 50 > $o/search_n2.json 2> $o/search_n2.err; echo "search rc=$?"; cat $o/search_n2.json; tail -3 $o/search_n2.err
python scratch/search_sharded.py 2 50 > $o/search_n1.json 2> $o/search_n1.err; cat $o/search_n1.json
