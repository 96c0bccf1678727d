#!/bin/bash
# GPU-box experiment: host topology, the packer alone, then KickEnv.step over the host pipelines (gpurun_out/pack_*.{txt,jsonl})
mkdir -p gpurun_out
{ nproc; lscpu | head -25; cat /sys/fs/cgroup/cpu.max 2>/dev/null; free -g | head -2; } > gpurun_out/pack_host.txt 2>&1
timeout 300 python -m pytest tests/test_task_parity_gpu.py tests/test_sibling_tasks_gpu.py -x -q -k "host" > gpurun_out/pack_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/pack_tests.log
: > gpurun_out/pack_alone.jsonl
for th in 8 15; do
  timeout 120 python tools/exp_host_pack.py --threads $th >> gpurun_out/pack_alone.jsonl 2>&1
done
timeout 600 python tools/exp_e2e.py --steps 48 --timeline --modes "${MODES:-staged_ce:4s,staged_pack:2,staged_pack:4,staged_pack:8}" > gpurun_out/pack_e2e.jsonl 2>&1; echo "e2e rc=$?"
cut -c1-1500 gpurun_out/pack_e2e.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/pack_alone.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d["workers"], d["spin_us"], d["pin"], d["best_total_ms"], d["runs"][-1])
PY
