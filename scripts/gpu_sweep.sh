# usage: bash scripts/gpu_sweep.sh TAG "sched codes" [pytest -k filter] [ncu kernel regex]
TAG=${1:-sweep}; SCHEDS=${2:-1045}; FILTER=${3:-}; NCUK=${4:-}
mkdir -p gpurun_out
if [ -n "$FILTER" ]; then timeout 1500 python -m pytest tests -m gpu -x -q -k "$FILTER" 2>&1 | tail -4; fi
for sched in $SCHEDS; do
  FEMTO_B200_COUNT_SCHED=$sched python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$sched.json 2> gpurun_out/${TAG}_$sched.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$sched.json"))
    r=d["roofline"]
    print("sched $sched", round(d["value"]/1e6,1), d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), "frac", r["frac"], "blocks", r["rank_blocks_distinct"], "ra", r["random_access"]["frac"], "hbm", d["config"]["engine"]["index_hbm_gib"], "loc", round(d["locate"]["value"]/1e6,1))
except Exception as e:
    print("sched $sched failed", e); print(open("gpurun_out/${TAG}_$sched.log").read()[-800:])
PY
done
if [ -n "$NCUK" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$NCUK -s 2 -c 1 -f -o gpurun_out/${TAG}_count python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
  ncu -i gpurun_out/${TAG}_count.ncu-rep --page details > gpurun_out/${TAG}_count_ncu_details.txt 2>&1
  grep -E "Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Eligible Warps|L1/TEX Cache Throughput" gpurun_out/${TAG}_count_ncu_details.txt
fi
