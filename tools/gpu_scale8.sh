#!/bin/bash
# Builder-side multi-GPU lines (one box, N ranks under torchrun): usage  bash tools/gpu_scale8.sh N
N=${1:-8}
mkdir -p gpurun_out; rm -f gpurun_out/scale${N}.jsonl
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-extras --no-cpu-baseline"
timeout 300 $T --steps 20 --warmup 5 >> gpurun_out/scale${N}.jsonl 2> gpurun_out/scale${N}.err || tail -3 gpurun_out/scale${N}.err
timeout 300 $T --workload xrb304 --zones 8192 --steps 3 --warmup 3 >> gpurun_out/scale${N}.jsonl 2> gpurun_out/scale${N}.err || tail -3 gpurun_out/scale${N}.err
timeout 300 $T --workload sn231_logft --zones 2048 --steps 3 --warmup 3 >> gpurun_out/scale${N}.jsonl 2> gpurun_out/scale${N}.err || tail -3 gpurun_out/scale${N}.err
python - <<PY
import json
for ln in open('gpurun_out/scale${N}.jsonl'):
    d=json.loads(ln)
    print(d['n_gpus'], d['config']['workload'][:50], 'zones/gpu', d['config']['zones_per_gpu'], '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'ms/step %.0f'%d['ms_per_step'], d['clocks'])
PY
