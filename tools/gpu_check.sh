#!/usr/bin/env bash
# Runs every -m gpu test file in its own process (a trapping kernel poisons only its own CUDA context), then
# smoke() and a short bench.  Logs land in gpurun_out/.  Usage: gpurun --timeout 1500 -- bash tools/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
summary=gpurun_out/check_summary.txt
: > "$summary"
for f in tests/test_gpu_*.py; do
  name=$(basename "$f" .py)
  timeout 420 python -m pytest "$f" -q -m gpu -x --no-header -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a "$summary"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$? $(tail -n 1 gpurun_out/smoke.log)" | tee -a "$summary"
if [ "${1:-}" != "quick" ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$? $(head -c 600 gpurun_out/bench.json)" | tee -a "$summary"
fi
cat "$summary"
