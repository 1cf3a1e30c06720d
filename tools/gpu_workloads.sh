mkdir -p gpurun_out; rm -f gpurun_out/workloads.jsonl
for a in "--workload xrb304 --zones 8192" "--workload sn231_logft --zones 16384" "--workload torch47_batch --zones 16384" "--workload alpha --zones 65536"; do
  timeout 600 python bench.py $a --steps 3 --warmup 3 --no-extras --no-cpu-baseline >> gpurun_out/workloads.jsonl 2>gpurun_out/wl.err || tail -3 gpurun_out/wl.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/workloads.jsonl'):
    d=json.loads(ln); r=d['roofline']
    print(d['config']['workload'][:60], '%.4g'%d['value'], 'ms/step %.0f'%d['ms_per_step'], d['config']['launch'], {k:v for k,v in r['phase_share'].items() if v}, r.get('structured',{}).get('zones_per_sm'), r.get('dense_fallbacks_per_lu'))
PY
