cd $GRAFT_REPO_ROOT
O=gpurun_out
for sp in 0 1 2; do TRACE=0 timeout 200 python profiles/tmem_trace.py AANET_MMA_SPIN=$sp > $O/spin_$sp.log 2>&1; done
paste $O/spin_0.log $O/spin_1.log $O/spin_2.log
for sp in 0 2; do AANET_B200_LIB=$PWD/aanet_b200/lib/libaanet_b200_prof.so timeout 200 python profiles/tmem_trace.py AANET_MMA_SPIN=$sp > $O/spin_prof_$sp.log 2>&1; done
