timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "paired" 2>&1 | tail -5
for cfg in "128 1045" "128 1044" "128 1046" "128 1024" "128 1025" "128 1013" "64 1025" "64 1026" "64 1024" "64 1014" "64 1015"; do
  set -- $cfg
  FEMTO_B200_COUNT_SCHED=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --block-bytes $1 --paired-levels 1 > gpurun_out/p_$1_$2.json 2> gpurun_out/p_$1_$2.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/p_$1_$2.json"))
    print("$1 $2", round(d["value"]/1e6,1), d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), "frac", d["roofline"]["frac"], "blocks", d["roofline"]["rank_blocks_distinct"], "hbm", d["config"]["index_hbm_gib"], "load", d["config"]["index_load_s"], "loc", round(d["locate"]["value"]/1e6,1))
except Exception as e:
    print("$1 $2 failed", e)
PY
done
