# One GPU box visit: parity suite, default bench, ncu captures of the count and walk kernels.
# (ncu serialises streams, so the profiled runs take the plain host-buffer path: FEMTO_B200_NO_STREAM)
# usage: bash scripts/gpu_round.sh <tag>   (files land in gpurun_out/<tag>_*)
TAG=${1:-round}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.log; echo bench_rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_default.json"))
e=d["config"]["engine"]
print("value", round(d["value"]/1e6,1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["ms_per_step"], "copy-only", d["e2e"]["copy_only_ms_per_step"], "e2e bytes", round(d["e2e_bytes"]["value"]/1e6,1), "load", e["index_load_s"], "hbm", e["index_hbm_gib"])
print("roofline", json.dumps(d["roofline"]))
print("cpu", json.dumps(d["cpu_baseline"])[:400])
print("parity", json.dumps(d["parity"])[:300])
print("locate", json.dumps(d["locate"])[:1500])
PY
FEMTO_B200_NO_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:count_sync -s 2 -c 1 -f -o gpurun_out/${TAG}_count python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-big-locate > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu -i gpurun_out/${TAG}_count.ncu-rep --page details > gpurun_out/${TAG}_count_ncu_details.txt 2>&1
FEMTO_B200_NO_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 12 -c 1 -f -o gpurun_out/${TAG}_walk python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_walk_bench.log 2>&1
ncu -i gpurun_out/${TAG}_walk.ncu-rep --page details > gpurun_out/${TAG}_walk_ncu_details.txt 2>&1
grep -E "walk_kernel|Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy" gpurun_out/${TAG}_walk_ncu_details.txt | head -12
FEMTO_B200_NO_STREAM=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"count_|walk_|occ_|probe_|clip|expand|total_" -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -4 gpurun_out/${TAG}_launches.csv
