#!/bin/bash
# Round 2, evidence trip (1 GPU): bench lines, launch list, ncu --set full of the hot kernels, full model, configs 3/5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/z_bench.json 2> $O/z_bench.err; echo "bench rc=$?" >> $O/z_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/z_bench_ref.json 2> $O/z_bench_ref.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/z_bench_c3.json 2> $O/z_bench_c3.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > $O/z_bench_c5.json 2> $O/z_bench_c5.err
timeout 600 python bench.py --config 5 --bf16-cost --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/z_bench_c5_bf16.json 2> $O/z_bench_c5_bf16.err
timeout 300 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/z_bench_b8.json 2> $O/z_bench_b8.err
W=/tmp/fm_w.pt
timeout 600 python profiles/full_model.py --variant stock --weights $W > $O/z_full_stock.json 2> $O/z_full_stock.err
timeout 600 python profiles/full_model.py --variant dropin --weights $W > $O/z_full_dropin.json 2> $O/z_full_dropin.err
# launch list of one eager step (cold, serialised) -- after the same command exited 0 without ncu
timeout 300 python profiles/profile_step.py --steps 3 > $O/z_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r02.csv python profiles/profile_step.py --steps 3 > $O/z_ncu1.log 2>&1
# full capture of the kernels of step 3
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"deform_tmem|conv_umma_kernel|corr_tma|csa_fuse|softargmin" -s 228 -c 84 -o /tmp/step_r02 python profiles/profile_step.py --steps 3 > $O/z_ncu2.log 2>&1
ncu -i /tmp/step_r02.ncu-rep --page raw --csv > $O/step_r02_raw.csv 2>/dev/null
ls -la /tmp/step_r02.ncu-rep $O/step_r02_raw.csv; du -sh $O
timeout 200 python profiles/timeline.py --graph > $O/timeline_graph_final.txt 2>/dev/null
TRACE=0 PIPE=1 timeout 200 python profiles/tmem_trace.py > $O/tmem_kernels_final.log 2>&1
AANET_B200_LIB=$PWD/aanet_b200/lib/libaanet_b200_base.so TRACE=0 PIPE=1 timeout 200 python profiles/tmem_trace.py > $O/tmem_kernels_a56f909.log 2>&1
head -c 500 $O/z_bench.json
