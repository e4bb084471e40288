# Quad-level layout: parity tests, then the count kernel at several register budgets.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multilevel or default_layout" 2>&1 | tail -5
for minb in 4 3 5 6 8; do
  FEMTO_B200_COUNT_SCHED=$((1020 + minb)) python bench.py --steps 5 --warmup 3 --no-cpu-baseline --levels 4 > gpurun_out/q_$minb.json 2> gpurun_out/q_$minb.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_$minb.json"))
    r=d["roofline"]
    print("quad minb $minb", round(d["value"]/1e6,1), d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), "frac", r["frac"], "blocks", r["rank_blocks_distinct"], "ra", r["random_access"], "hbm", d["config"]["index_hbm_gib"], "load", d["config"]["index_load_s"], "loc", round(d["locate"]["value"]/1e6,1))
except Exception as e:
    print("quad $minb failed", e); print(open("gpurun_out/q_$minb.log").read()[-800:])
PY
done
