#!/bin/bash
# compute-sanitizer racecheck + memcheck of the count, clip/scan/expand and walk kernels through the two
# C tools (no Python): femto_request_b200 (count batches of 522 patterns) and femto_multiquery_b200 -locate,
# on the committed golden index mixed_1500.
mkdir -p gpurun_out
IDX=tests/golden/mixed_1500/index
printf '# number=8 length=2 file=synthetic forbidden=\nACGTCATGaaAAGGCC' > /tmp/p.pc
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck memcheck; do
  timeout 10 $CS --tool $tool --error-exitcode 9 oracle/_ref/femto_multiquery_b200 $IDX -locate 100000 < /tmp/p.pc \
    > gpurun_out/r02_sanitizer_${tool}_locate.txt 2>&1; echo "$tool locate rc=$?"
  timeout 10 $CS --tool $tool --error-exitcode 9 oracle/_ref/femto_request_b200 $IDX "string_rows_all 71" \
    > gpurun_out/r02_sanitizer_${tool}_count.txt 2>&1; echo "$tool count rc=$?"
done
tail -n 4 gpurun_out/r02_sanitizer_*.txt
