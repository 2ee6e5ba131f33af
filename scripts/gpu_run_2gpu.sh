#!/bin/bash
# scratch: GPU run 21 (2 GPUs) - final verification: gpu tests, smoke, 2-rank merge check, 2-GPU bench
mkdir -p gpurun_out
T=gpurun_out/run21
timeout 400 python -m pytest tests -m gpu -x -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > ${T}_dist_check.log 2>&1
echo "dist check exit $?"; tail -2 ${T}_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > ${T}_bench_2gpu.log 2>&1
grep -h '^{"metric' ${T}_bench_2gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['volume_merge'])" || tail -5 ${T}_bench_2gpu.log
