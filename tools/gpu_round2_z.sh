#!/bin/bash
# the other presets on the shipped build (BASELINE configs[2] = pare_w32 at batch 128; pare_r50 and cliff_w48cls at batch 256)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/r02h_bench_presets.jsonl
for cfg in "pare_w32 128" "pare_r50 256" "cliff_w48cls 256"; do
  set -- $cfg
  timeout 300 python bench.py --preset $1 --batch $2 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_preset.err | grep '^{' >> gpurun_out/r02h_bench_presets.jsonl; echo "$1 rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r02h_bench_presets.jsonl'):
    d = json.loads(l)
    print(d['config']['preset'], d['config']['crops_per_gpu'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_err'],
          'split', (d.get('other_precision_mode') or {}).get('value'), (d.get('other_precision_mode') or {}).get('parity_err'))
PY
