# whole-step A/B: bench.py with the a56f909 library and with the tree's
cd $GRAFT_REPO_ROOT
O=gpurun_out
AANET_B200_LIB=$PWD/aanet_b200/lib/libaanet_b200_base.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/abb_base.json 2> $O/abb_base.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/abb_cur.json 2> $O/abb_cur.err
python - <<'PY'
import json
for n in ("base","cur"):
    try:
        d=json.loads(open("gpurun_out/abb_%s.json"%n).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(n, "pairs/s %.1f  ms %.4f  dcn launch %.1f us  frac %.3f" % (d["value"], d["ms_per_step"], r["us_per_launch"], r["frac"]), {k:round(v["us"],1) for k,v in r["other_kernels"].items()})
    except Exception as e:
        print(n, "ERR", e, open("gpurun_out/abb_%s.err"%n).read()[-500:])
PY
