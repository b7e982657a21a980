"""Column-sharded commit (lambdaworks_cairo_prover_b200/distributed.py): world_size-2 and -4 gloo runs on CPU
for the sharding logic, and an NCCL run on real GPUs when at least two are visible."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from lambdaworks_cairo_prover_b200.distributed import build_top, column_shards

WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(mode, world, logn, n_cols, blowup, groups=1, exchange="p2p", timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), WORKER, mode, str(logn), str(n_cols), str(blowup),
           str(groups), exchange]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_OK" in res.stdout
    return res.stdout


def test_column_shards_match_the_survey_split():
    assert [b - a for a, b in column_shards(33, 8)] == [5, 4, 4, 4, 4, 4, 4, 4]
    assert column_shards(34, 2) == [(0, 17), (17, 34)]
    assert [b - a for a, b in column_shards(2, 2)] == [1, 1]
    assert sum(b - a for a, b in column_shards(52, 4)) == 52


def test_build_top_is_the_heap_rule():
    from oracle import pyoracle as O
    leaves = [bytes([i]) * 32 for i in range(4)]
    top = build_top(leaves, O.keccak256)
    assert top[3:] == leaves
    assert top[1] == O.keccak256(leaves[0] + leaves[1]) and top[2] == O.keccak256(leaves[2] + leaves[3])
    assert top[0] == O.keccak256(top[1] + top[2])


@pytest.mark.parametrize("world,logn,n_cols,blowup,groups", [(2, 6, 5, 4, 1), (4, 5, 7, 2, 1), (2, 4, 2, 8, 1), (2, 5, 9, 4, 3),
                                                             (4, 4, 10, 2, 2)])
def test_sharded_commit_gloo(world, logn, n_cols, blowup, groups):
    launch("gloo", world, logn, n_cols, blowup, groups)


def test_sharded_commit_gloo_packed_all_to_all():
    launch("gloo", 2, 5, 7, 4, 1, "a2a")
    launch("gloo", 4, 4, 9, 2, 1, "a2a")


@pytest.mark.gpu
def test_sharded_commit_nccl():
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least two GPUs")
    world = 4 if g >= 4 else 2
    launch("nccl", world, 12, 33, 8)
    launch("nccl", 2, 10, 3, 4)
    launch("nccl", world, 12, 33, 8, groups=2)
    launch("nccl", world, 11, 18, 4, exchange="a2a")


@pytest.mark.parametrize("world,args", [(2, ("3", "4", "3", "3", "1")), (4, ("10", "2", "4", "5", "3")), (1, ("3", "4", "2", "3", "1")),
                                        (2, ("20", "8", "5", "7", "2")), (8, ("20", "4", "3", "3", "2"))])
def test_sharded_cairo_proof_gloo(world, args):
    """The orchestration of the sharded Cairo prover (column shards, exchange, halos, gathers, row-block trees of the trace tables,
    of (H1, H2) and of the FRI layers, the pairwise fold exchange, the collapse to one rank, grinding split over the ranks, transcript
    replay, openings served by the row owners, serialization) on CPU ranks, with the GPU backend replaced by its oracle double
    (tests/cairo_oracle_backend.py): bytes == the oracle's single-process prover, for both exchange kinds and with the FRI layers
    sharded down to 8 / 32 evaluations or collapsed at once."""
    worker = os.path.join(ROOT, "tests", "dist_cairo_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), worker] + list(args) + ["gloo"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_CAIRO_OK" in res.stdout


@pytest.mark.gpu
def test_sharded_cairo_proof_nccl():
    """ONE Cairo proof over 2 (and 4) GPUs == the single-GPU proof, byte for byte."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least two GPUs")
    worker = os.path.join(ROOT, "tests", "dist_cairo_worker.py")
    for world, args in ((2, ("100", "4", "3", "3", "1")), (4 if g >= 4 else 2, ("1000", "8", "5", "7", "6"))):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), worker] + list(args)
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
        assert "DIST_CAIRO_OK" in res.stdout


@pytest.mark.parametrize("world,args", [(2, ("6", "4", "2")), (4, ("7", "2", "3")), (2, ("5", "8", "2")), (4, ("8", "2", "2")), (8, ("9", "4", "3"))])
def test_sharded_single_column_gloo(world, args):
    """ONE column over several ranks -- four-step geometry, the transposes and redistributions, the row-block tree and
    the openings on CPU ranks (the transform phases restated in python integers) == the oracle's interpolate_and_commit."""
    worker = os.path.join(ROOT, "tests", "dist_column_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), worker, "gloo"] + list(args)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_COLUMN_OK" in res.stdout


@pytest.mark.gpu
def test_sharded_single_column_nccl():
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least two GPUs")
    worker = os.path.join(ROOT, "tests", "dist_column_worker.py")
    for world, args in ((2, ("14", "4")), (4 if g >= 4 else 2, ("16", "8")), (g if g in (2, 4, 8) else 2, ("20", "4"))):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), worker, "nccl"] + list(args)
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
        assert "DIST_COLUMN_OK" in res.stdout


def launch_native(world, *args, worker="dist_native_worker.py", ok="NATIVE_SHARDED_OK"):
    """One plain python process per GPU running a tests/dist_native_*worker.py (no torch.distributed: the library calls NCCL itself).
    args: the worker's arguments between `world` and the id file (dist_native_worker: logn n_cols blowup groups [mem])."""
    import tempfile
    worker = os.path.join(ROOT, "tests", worker)
    args = [str(a) for a in args]
    tail = [args.pop()] if args and args[-1] in ("host", "device") else []
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "nccl_id")
        procs = [subprocess.Popen([sys.executable, worker, str(r), str(world)] + args + [idfile] + tail,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
                 for r in range(world)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=600)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(out)
        for r, (p, out) in enumerate(zip(procs, outs)):
            assert p.returncode == 0, "rank %d:\n%s" % (r, out[-4000:])
            assert ok in out


@pytest.mark.gpu
def test_native_sharded_commit_one_rank():
    """s252_interpolate_and_commit_sharded / s252_sharded_commit_open with a communicator of ONE rank (runs on a single-GPU box):
    the whole C++ path -- NCCL bound at run time, group bookkeeping, in-place subtree, top levels, packed openings -- against
    the oracle's root, rows and paths."""
    launch_native(1, 10, 5, 4, 2)
    launch_native(1, 8, 3, 8, 3, "device")


@pytest.mark.gpu
def test_native_sharded_commit_nccl():
    """The same over 2 (and 4) GPUs: column shards of unequal width, more pipeline groups than some ranks have columns."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least two GPUs")
    launch_native(2, 10, 5, 4, 2)
    launch_native(2, 9, 3, 8, 4, "device")
    if g >= 4:
        launch_native(4, 12, 33, 8, 3)


@pytest.mark.gpu
def test_native_sharded_cairo_proof_one_rank():
    """s252_cairo_prove_sharded with a communicator of ONE rank (runs on a single-GPU box): the whole C++ orchestration -- row-block
    building blocks, halos, the packed openings, serialization -- must reproduce s252_cairo_prove's bytes."""
    launch_native(1, 100, 4, 3, 3, 1, worker="dist_native_cairo_worker.py", ok="NATIVE_CAIRO_OK")


@pytest.mark.gpu
def test_native_sharded_cairo_proof_nccl():
    """ONE Cairo proof over 2 (and 4) GPUs through the single collective C-ABI call == the single-GPU proof, byte for byte; FRI layers
    sharded down to 32 evaluations, grinding split over the ranks."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least two GPUs")
    launch_native(2, 100, 4, 3, 3, 1, worker="dist_native_cairo_worker.py", ok="NATIVE_CAIRO_OK")
    launch_native(4 if g >= 4 else 2, 1000, 8, 5, 7, 6, worker="dist_native_cairo_worker.py", ok="NATIVE_CAIRO_OK")


def test_native_and_python_paths_shard_columns_alike():
    """The C-ABI collective calls (sharded.py / csrc/sharded.cuh: shard_range) and the torch.distributed path (distributed.py:
    column_shards) must cut a table into the same contiguous column ranges, and pipeline groups by the same rule."""
    from lambdaworks_cairo_prover_b200 import distributed as D, sharded as S
    for n_cols in (1, 2, 3, 18, 33, 34, 43, 52):
        for world in (1, 2, 4, 8, 16):
            if n_cols < world:
                continue
            assert [S.my_columns(n_cols, world, r) for r in range(world)] == D.column_shards(n_cols, world)
    for width in (1, 4, 5, 9):
        for groups in (1, 2, 3, 4):
            g = max(1, min(groups, width))
            assert [S.my_columns(width, g, k) for k in range(g)] == D.group_ranges(width, groups)
