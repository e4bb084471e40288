# One GPU box visit: parity suite, default bench, probe table, ncu captures of the count kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/default.json 2> gpurun_out/default.log; echo bench_rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/default.json"))
print("value", round(d["value"]/1e6,1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), d["e2e"]["ms_per_step"])
print("roofline", json.dumps(d["roofline"]))
print("cpu", json.dumps(d["cpu_baseline"])[:400])
print("parity", json.dumps(d["parity"])[:300])
print("locate", json.dumps(d["locate"])[:500])
PY
python - <<'PY'
import femto_b200 as fb, glob, json
path = sorted(glob.glob("/tmp/femto_b200_cache/*"))[0]
ix = fb.Index(path)
out = {}
for b in (32, 64, 128):
    r = ix.probe_random_reads(b, steps=400)
    out[b] = r
    print("probe", b, "B:", round(r["accesses_per_s"]/1e9, 2), "G/s", round(r["gb_per_s"], 1), "GB/s")
json.dump(out, open("gpurun_out/probe.json", "w"))
PY
ncu --set full --clock-control none --import-source on -k regex:count_sync -s 2 -c 1 -f -o gpurun_out/r01_count_paired64 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu -i gpurun_out/r01_count_paired64.ncu-rep --page details > gpurun_out/r01_count_paired64_ncu_details.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"count_|walk_|occ_|probe_" -c 60 --csv --log-file gpurun_out/r01_launches_paired64.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -5 gpurun_out/r01_launches_paired64.csv
