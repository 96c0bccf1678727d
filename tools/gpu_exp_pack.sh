#!/bin/bash
# GPU-box experiment: KickEnv.step over the staged_pack host pipeline, chunk-size schedules (gpurun_out/pack_e2e.jsonl)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_task_parity_gpu.py tests/test_sibling_tasks_gpu.py -x -q -k "host" > gpurun_out/pack_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/pack_tests.log
: > gpurun_out/pack_e2e.jsonl
timeout 600 python tools/exp_e2e.py --steps 64 --modes "${MODES:-staged_pack:4,staged_pack:4,staged_ce:4s}" >> gpurun_out/pack_e2e.jsonl 2>&1; echo "e2e rc=$?"
cut -c1-300 gpurun_out/pack_e2e.jsonl
