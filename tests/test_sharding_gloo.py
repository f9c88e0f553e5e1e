"""N > 1 path on CPU: world_size-2 gloo process groups exercising the candidate and k-slab sharding
logic of optimize_stencil_b200/sharding.py with the oracle injected as the local evaluator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import np_oracle as O
from optimize_stencil_b200 import sharding

XC = np.array([0.5, .02, .03, .01, .02, .03, .01, -.02, -.03, -.01])
N = 12


def coeff_batch(B):
    rng = np.random.default_rng(0)
    P = XC + 1e-3 * rng.standard_normal((B, 10))
    P[B // 2] = [0.4, .6, -.5, .4, .3, -.7, .2, .24, -.9, .1]   # negative sqrtarg -> NaN sum
    return np.stack([O.coefficients("StencilFree3D", p) for p in P])


def oracle_evaluator(block, x_begin, x_end, x_stride=1):
    out = np.empty((4, block.shape[0]))
    grid = O.Grid(3, N, xslab=(x_begin, x_end, x_stride))
    for b in range(block.shape[0]):
        S, amin, amax, nn = O.device_contract(grid, block[b].numpy())
        out[:, b] = (S, amin if amin < 0 else 0.0, max(amax, 0.0), nn)
    return torch.from_numpy(out)


def _worker(rank, world, port, mode, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = sharding.ShardedObjective(oracle_evaluator, nx=N, mode=mode)
        assert (sh.rank, sh.world) == (rank, world)
        out = sh.evaluate(torch.from_numpy(coeff_batch(B)))
        S, Amin, Amax, nan = sharding.host_evaluate(sh, coeff_batch(B), torch.device("cpu"))
        assert np.array_equal(S, out[0].numpy(), equal_nan=True) and nan.dtype == np.int64
        q.put((rank, out.numpy()))
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode,B", [("candidates", 7), ("candidates", 2), ("candidates", 8), ("slabs", 5),
                                    ("slabs-contiguous", 5)])
def test_world_size_2(mode, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    assert all(not isinstance(v, str) for v in got.values()), got
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = oracle_evaluator(torch.from_numpy(coeff_batch(B)), 0, N).numpy()
    for r in range(2):
        assert got[r].shape == (4, B)
        np.testing.assert_array_equal(got[r][1:], whole[1:])          # extrema and NaN counts: exact
        if mode == "candidates":
            np.testing.assert_array_equal(got[r][0], whole[0])        # same evaluator, same order
        else:
            ok = ~np.isnan(whole[0])
            assert np.array_equal(np.isnan(got[r][0]), ~ok)
            assert np.max(np.abs(got[r][0][ok] - whole[0][ok]) / np.abs(whole[0][ok])) < 1e-13
    np.testing.assert_array_equal(got[0], got[1])                     # every rank holds the same answer


def test_split_range_partitions():
    for n, w in [(4096, 8), (7, 2), (5, 8), (2048, 3), (1, 1)]:
        blocks = [sharding.split_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        sizes = sharding.block_sizes(n, w)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def test_combine_slab_partials_propagates_nan_and_order():
    parts = torch.tensor([[[1.0, 2.0], [-0.5, 0.0], [3.0, 1.0], [0.0, 1.0]],
                          [[0.5, float("nan")], [0.0, float("nan")], [4.0, 2.0], [2.0, 0.0]]], dtype=torch.float64)
    out = sharding.combine_slab_partials(parts).numpy()
    assert out[0][0] == 1.5 and np.isnan(out[0][1])
    assert out[1][0] == -0.5 and np.isnan(out[1][1]) and np.isnan(out[2][1])
    assert out[2][0] == 4.0 and out[3][0] == 2.0 and out[3][1] == 1.0
    single = sharding.ShardedObjective(oracle_evaluator, nx=N, mode="slabs")
    assert (single.rank, single.world) == (0, 1) and single.my_planes() == (0, N, 1)
    assert sharding.ShardedObjective(oracle_evaluator, nx=N, mode="slabs-contiguous").my_planes() == (0, N, 1)
