#!/bin/bash
# compute-sanitizer over smoke() (replay + one SAC1 update incl. the tcgen05 / TMA kernels) and over every replay kernel
# family (GPU box, via gpurun); per-tool logs in gpurun_out/, one-line summaries in gpurun_out/r02_sanitizer_summary.txt
mkdir -p gpurun_out
S=gpurun_out/r02_sanitizer_summary.txt
echo "compute-sanitizer $(compute-sanitizer --version | tail -1) on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)" > $S
for tool in memcheck racecheck synccheck initcheck; do
  for job in smoke replay; do
    if [ $job = smoke ]; then cmd=(python -c "import __graft_entry__ as g; g.smoke()"); else cmd=(python tools/sanitize_replay.py); fi
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 "${cmd[@]}" > gpurun_out/san_${tool}_${job}.log 2>&1
    echo "$tool / $job: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_${job}.log | tail -1) ; $(grep -cE 'smoke ok|sanitize_replay done' gpurun_out/san_${tool}_${job}.log) completion line(s)" >> $S
  done
done
cat $S
