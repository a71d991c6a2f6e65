#!/bin/bash
# Round 2, multi-GPU trip: bash profiles/run_r02_multi.sh N   (N = 2, 4 or 8 GPUs of one box)
# configs 2 / 3 / 5 (+ bf16 cost volume) batch-sharded over N ranks, the 2-GPU DDP training step, the 2-GPU test.
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/m${N}_bench_c2.json 2> $O/m${N}_bench_c2.err
timeout 600 $RUN bench.py --gpus $N --config 3 --steps 5 --warmup 3 > $O/m${N}_bench_c3.json 2> $O/m${N}_bench_c3.err
timeout 600 $RUN bench.py --gpus $N --config 5 --steps 3 --warmup 3 > $O/m${N}_bench_c5.json 2> $O/m${N}_bench_c5.err
timeout 600 $RUN bench.py --gpus $N --config 5 --bf16-cost --steps 3 --warmup 3 --no-e2e > $O/m${N}_bench_c5_bf16.json 2> $O/m${N}_bench_c5_bf16.err
if [ "$N" = "2" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/ddp_train_step.py > $O/m2_ddp.json 2> $O/m2_ddp.err; echo "ddp rc=$?" >> $O/m2_ddp.err
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -k second_device > $O/m2_pytest.log 2>&1
  tail -3 $O/m2_pytest.log; tail -2 $O/m2_ddp.err; cat $O/m2_ddp.json
fi
for f in $O/m${N}_bench_c*.json; do head -c 220 $f; echo; done
