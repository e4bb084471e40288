# Lane-per-pattern quad count kernel: parity tests, register-budget sweep, ncu capture.
TAG=${1:-lane}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for sched in 1014 1013 1015 1025; do
  FEMTO_B200_COUNT_SCHED=$sched python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$sched.json 2> gpurun_out/${TAG}_$sched.log
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$sched.json"))
    r=d["roofline"]
    print("sched $sched", round(d["value"]/1e6,1), d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), "frac", r["frac"], "blocks", r["rank_blocks_distinct"], "ra", r["random_access"]["frac"], "hbm", d["config"]["index_hbm_gib"], "loc", round(d["locate"]["value"]/1e6,1))
except Exception as e:
    print("sched $sched failed", e); print(open("gpurun_out/${TAG}_$sched.log").read()[-800:])
PY
done
ncu --set full --clock-control none --import-source on -k regex:count_quad_lane -s 2 -c 1 -f -o gpurun_out/${TAG}_count python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu -i gpurun_out/${TAG}_count.ncu-rep --page details > gpurun_out/${TAG}_count_ncu_details.txt 2>&1
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Eligible Warps|L1/TEX Cache Throughput" gpurun_out/${TAG}_count_ncu_details.txt
