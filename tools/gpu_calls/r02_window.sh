#!/bin/bash
# --order window: rank-sorted rows, whole-segment chunks in column-window order (plan only, same kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -x -q --tb=short -p no:cacheprovider -k window > gpurun_out/r02_pytest_window.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_window.log; tail -4 gpurun_out/r02_pytest_window.log
: > gpurun_out/r02_window.jsonl
for v in "--order degree" "--order window --window-key mid" "--order window --window-key first" "--order window --window-key mid --window-wide" "--order window --window-key last"; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-parity $v > gpurun_out/r02_window_one.log 2>&1
  grep '^{' gpurun_out/r02_window_one.log >> gpurun_out/r02_window.jsonl || tail -5 gpurun_out/r02_window_one.log
  python - "$v" <<'PY'
import json, sys
l = [x for x in open("gpurun_out/r02_window_one.log") if x.startswith("{")]
if l:
    d = json.loads(l[-1]); print(sys.argv[1], "ms/pass", round(d["ms_per_step"], 3), "launch_ms", round(d["roofline"]["launch_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "build_s", d["config"]["graph_build_s"])
PY
done
# counters of the best-looking variant
timeout 300 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --clock-control none -k regex:spmm_stream_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/r02_prof_window -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras --order window --window-key mid > gpurun_out/r02_ncu_window.log 2>&1
ls -la gpurun_out/r02_prof_window.ncu-rep
