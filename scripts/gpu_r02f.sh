# round-2 single-GPU visit: parity suite, default bench (+ pointer API), ncu captures, streamed-call stress.
# every step under its own timeout
mkdir -p gpurun_out
T=r02f
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 --pointer-api > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.log; echo bench_rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_default.json"))
print("value", round(d["value"]/1e6,1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["ms_per_step"], "copy-only", d["e2e"]["copy_only_ms_per_step"], "e2e bytes", round(d["e2e_bytes"]["value"]/1e6,1), "ptr", d["e2e_pointer_api"])
l=d["locate"]; print("locate", l["value"], l["ms_per_batch"], l["roofline"]["kernel_ms"], l["roofline"]["random_access"], "big", l["whole_batch"]["value"], l["whole_batch"]["ms_per_batch"], l["whole_batch"]["roofline"]["kernel_ms"], l["whole_batch"]["roofline"]["random_access"])
PY
FEMTO_B200_NO_STREAM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:count_sync -s 2 -c 1 -f -o gpurun_out/${T}_count python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-big-locate > gpurun_out/${T}_ncu_bench.log 2>&1
ncu -i gpurun_out/${T}_count.ncu-rep --page details > gpurun_out/${T}_count_ncu_details.txt 2>&1
FEMTO_B200_NO_STREAM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 12 -c 1 -f -o gpurun_out/${T}_walk python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_walk_bench.log 2>&1
ncu -i gpurun_out/${T}_walk.ncu-rep --page details > gpurun_out/${T}_walk_ncu_details.txt 2>&1
grep -E "walk_kernel|  Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Executed Ipc Active" gpurun_out/${T}_walk_ncu_details.txt | head -12
FEMTO_B200_NO_STREAM=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"count_|walk_|occ_|probe_|clip|expand|total_" -c 80 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/${T}_launches.csv | cut -c1-300
timeout 300 python scripts/stream_stress.py 1500 7 2>&1 | tail -2 | tee gpurun_out/${T}_stream_stress.txt
