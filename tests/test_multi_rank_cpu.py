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


TS_WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np
import torch.distributed as dist
import bench
import xnet_b200 as X
from conftest import DATA, PROBLEMS, control_text
from oracle import xnet_oracle as O
rank, world, local = bench.dist_setup(2)
txt = control_text(*PROBLEMS['alpha'])
ctl = X.read_controls(txt)
net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
Z = 3                                              # zones per rank; global zone z uses rho_scale(z): distinct outputs
gz, zf = bench.zone_block(rank, world, Z, 1)
scale = 1.0 + 1e-3 * np.linspace(-1, 1, world * Z)
h = O.Problem.from_control(txt, DATA).run(nzone=Z, hist_cap=200, rho_scale=scale[gz], flux=True, ctl_override=dict(kstmx=150))
for k, z in enumerate(gz):                        # ts file name = base + zero-padded GLOBAL zone number (net.F90:283-285)
    X.write_ts(X.name_ordered(os.path.join(%(out)r, 'ts_'), int(z) + 1, world * Z), net, ctl, None, h, zone=k)
bench.barrier(world)
if rank == 0:
    print(json.dumps({'files': sorted(os.listdir(%(out)r))}))
dist.destroy_process_group()
"""


def test_gloo_world2_ts_files_per_global_zone(tmp_path, built):
    """Results are gathered at output: each rank writes the ts files of its own zone block, named by GLOBAL zone
    number; two ranks together produce exactly the files (names and bytes) that one process produces for all zones."""
    import xnet_b200 as X
    from conftest import DATA, PROBLEMS, control_text
    from oracle import xnet_oracle as O
    out = tmp_path / 'ranks'; out.mkdir()
    script = tmp_path / 'ts_worker.py'
    script.write_text(TS_WORKER % dict(root=ROOT, out=str(out)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(line) == 1 and json.loads(line[0])['files'] == ['ts_%d' % k for k in range(1, 7)]
    txt = control_text(*PROBLEMS['alpha'])
    ctl = X.read_controls(txt); net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
    scale = 1.0 + 1e-3 * np.linspace(-1, 1, 6)
    h = O.Problem.from_control(txt, DATA).run(nzone=6, hist_cap=200, rho_scale=scale, flux=True, ctl_override=dict(kstmx=150))
    single = tmp_path / 'single'; single.mkdir()
    for z in range(6):
        X.write_ts(X.name_ordered(str(single / 'ts_'), z + 1, 6), net, ctl, None, h, zone=z)
        assert (single / ('ts_%d' % (z + 1))).read_bytes() == (out / ('ts_%d' % (z + 1))).read_bytes()
    assert (out / 'ts_1').read_bytes() != (out / 'ts_6').read_bytes()          # the zones really differ
