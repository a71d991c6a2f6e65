#!/bin/bash
# Round 2, GPU trip A: parity suite, bench (configs 2/3/5), full model stock vs drop-in, probes, sanitizer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/a_gpu.txt 2>&1
nproc >> $O/a_gpu.txt
timeout 120 profiles/probes/tma_umma_probe > $O/a_probe.log 2>&1; echo "probe rc=$?" >> $O/a_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/a_bench.json 2> $O/a_bench.err; echo "bench rc=$?" >> $O/a_bench.err
W=/tmp/fm_w.pt
timeout 600 python profiles/full_model.py --variant stock --weights $W --out /tmp/stock.npz > $O/a_full_stock.json 2> $O/a_full_stock.err
timeout 600 python profiles/full_model.py --variant dropin --weights $W --out /tmp/dropin.npz > $O/a_full_dropin.json 2> $O/a_full_dropin.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/a_bench_c3.json 2> $O/a_bench_c3.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > $O/a_bench_c5.json 2> $O/a_bench_c5.err
timeout 600 python bench.py --config 5 --bf16-cost --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/a_bench_c5_bf16.json 2> $O/a_bench_c5_bf16.err
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_probe.py > $O/a_memcheck.log 2>&1; echo "rc=$?" >> $O/a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python profiles/sanitize_probe.py > $O/a_racecheck.log 2>&1; echo "rc=$?" >> $O/a_racecheck.log
tail -3 $O/a_pytest.log
head -c 600 $O/a_bench.json
