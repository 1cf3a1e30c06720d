"""world_size-2 tests of the multi-rank host logic on CPU (gloo): the zone block partition, the only reduction
the job performs (max time, summed Newton iterations) and the launch contract of the reference arm under
torchrun (rank 0 prints the single JSON line, the other ranks exit 0)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def test_zone_block_partition_is_disjoint_and_balanced():
    import bench
    for world in (1, 2, 4, 8):
        for K in (1, 4, 6):
            seen = []
            mixes = []
            for r in range(world):
                gz, zf = bench.zone_block(r, world, 48, K)
                seen.append(gz); mixes.append(np.bincount(zf, minlength=K))
                assert np.array_equal(zf, gz % K)
            allz = np.concatenate(seen)
            assert np.array_equal(np.sort(allz), np.arange(world * 48))        # disjoint, complete
            if 48 % K == 0:
                assert all(np.array_equal(m, mixes[0]) for m in mixes)         # every rank gets the same work mix


WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
rank, world, local = bench.dist_setup(2)
assert world == 2 and dist.get_backend() == 'gloo'
gz, zf = bench.zone_block(rank, world, 5, 4)
t, n = bench.reduce_max_sum(world, 1.0 + rank, float(gz.sum()))
bench.barrier(world)
if rank == 0:
    print(json.dumps({'tmax': t, 'nsum': n}))
dist.destroy_process_group()
"""


def test_gloo_world2_reduction(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(line) == 1
    d = json.loads(line[0])
    assert d['tmax'] == 2.0 and d['nsum'] == float(sum(range(10)))


def test_reference_arm_under_torchrun_prints_one_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'bench.py'),
                          '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--workload', 'alpha',
                          '--cpu-zones', '4'], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['e2e']['h2d_bytes_per_step'] == 0
