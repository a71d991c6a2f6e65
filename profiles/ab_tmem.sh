# A/B of the TMEM-A kernels: the library of commit a56f909 (lib/libaanet_b200_base.so, built by hand) against the tree's
cd $GRAFT_REPO_ROOT
O=gpurun_out
AANET_B200_LIB=$PWD/aanet_b200/lib/libaanet_b200_base.so TRACE=0 timeout 200 python profiles/tmem_trace.py > $O/ab_base.log 2>&1
TRACE=0 timeout 200 python profiles/tmem_trace.py > $O/ab_cur.log 2>&1
paste $O/ab_base.log $O/ab_cur.log
timeout 500 python -m pytest tests -m gpu -x -q -k "deform or tail or dense or tmem or hot_path or fused or mdcn or reference_op" > $O/ab_pytest.log 2>&1; tail -3 $O/ab_pytest.log
