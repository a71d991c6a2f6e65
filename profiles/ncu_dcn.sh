# ncu --set full + source page of the 1/3-scale DCN launch (small offsets) -> gpurun_out/dcn_source.csv, dcn_raw.csv
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python profiles/ncu_deform.py 0.1 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:deform_tmem -s 1 -c 1 -o /tmp/dcn_cur python profiles/ncu_deform.py 0.1 > $O/ncu_dcn.log 2>&1
ncu -i /tmp/dcn_cur.ncu-rep --page source --csv > $O/dcn_source.csv 2>/dev/null
ncu -i /tmp/dcn_cur.ncu-rep --page raw --csv > $O/dcn_raw.csv 2>/dev/null
ls -la $O/dcn_source.csv
