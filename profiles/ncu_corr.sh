# ncu --set full (+ source page) of the TMA-staged correlation at the three scales of config 2
cd $GRAFT_REPO_ROOT
O=gpurun_out
sed -i 's/ops.correlation(L, R, D)/ops.correlation_nhwc(L, R, D)/' profiles/ncu_corr.py
timeout 200 python profiles/ncu_corr.py > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_tma -c 6 -o /tmp/corr_r02 python profiles/ncu_corr.py > $O/ncu_corr.log 2>&1
ncu -i /tmp/corr_r02.ncu-rep --page raw --csv > $O/corr_r02_raw.csv 2>/dev/null
ncu -i /tmp/corr_r02.ncu-rep --page source --csv --kernel-name regex:corr_tma --launch-skip 1 --launch-count 1 > $O/corr_r02_source.csv 2>/dev/null
ls -la $O/corr_r02_raw.csv $O/corr_r02_source.csv
