#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 &&
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?"
tail -5 gpurun_out/sanitizer_memcheck.log
