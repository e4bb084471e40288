mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --pointer-api > gpurun_out/r02e_default.json 2> gpurun_out/r02e_default.log; echo bench_rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02e_default.json"))
e=d["config"]["engine"]
print("value", round(d["value"]/1e6,1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["ms_per_step"], "copy-only", d["e2e"]["copy_only_ms_per_step"], "e2e bytes", round(d["e2e_bytes"]["value"]/1e6,1), d["e2e_bytes"]["ms_per_step"], d["e2e_bytes"]["copy_only_ms_per_step"], "ptr", d["e2e_pointer_api"])
print("roofline", json.dumps(d["roofline"])[:600])
print("locate", json.dumps(d["locate"])[:2200])
PY
FEMTO_B200_NO_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:count_sync -s 2 -c 1 -f -o gpurun_out/r02e_count python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-big-locate > gpurun_out/r02e_ncu_bench.log 2>&1
ncu -i gpurun_out/r02e_count.ncu-rep --page details > gpurun_out/r02e_count_ncu_details.txt 2>&1
FEMTO_B200_NO_STREAM=1 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 10 -c 1 -f -o gpurun_out/r02e_walk python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02e_ncu_walk_bench.log 2>&1
ncu -i gpurun_out/r02e_walk.ncu-rep --page details > gpurun_out/r02e_walk_ncu_details.txt 2>&1
grep -E "walk_kernel|  Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy" gpurun_out/r02e_walk_ncu_details.txt | head -12
FEMTO_B200_NO_STREAM=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"count_|walk_|occ_|probe_|clip|expand|total_" -c 80 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r02e_launches.csv | cut -c1-300
echo ==== config3 16 GiB
python bench.py --kind english --corpus-mib 16384 --patterns zipf --steps 10 --warmup 3 > gpurun_out/r02e_config3_16gib.json 2> gpurun_out/r02e_config3_16gib.log; echo rc=$?
tail -c 3000 gpurun_out/r02e_config3_16gib.json; grep -v "build:" gpurun_out/r02e_config3_16gib.log | tail -6
