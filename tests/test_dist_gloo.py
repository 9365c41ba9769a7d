"""world_size-2 gloo test (CPU) of the multi-GPU host logic: pair-range sharding + all-reduce of
the kappa sums reproduces the single-process result.  The rank-local compute is the oracle here
(no GPU in this container); on the GPU box the same code path runs with RPMD.recross_children."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from caracal_b200.shard import recross_sharded, shard_range, umbrella_sharded  # noqa: E402


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 512, 1000):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == n
            pos = 0
            for s, c in blocks:
                assert s == pos
                pos += c
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


NPAIRS, EVOL, XI = 6, 25, 0.98


def _compute_factory():
    from oracle import oracle as O
    from tests import common as C
    name, nb = "h3", 4
    o = O.System(name, nb, C.masses(name), C.beta_calc_rate(300.0), C.dt_au(0.1))
    o.set_mechanism(C.mechanism(name))
    qp = np.array([C.ring_polymer(name, nb, np.random.default_rng(k), 0.02) for k in range(2)])

    def compute(pair0, npairs):
        if npairs == 0:
            return np.zeros(EVOL), 0.0
        num, den, st = o.recross_children(qp, pair0, npairs, EVOL, XI, C.SEED, nthreads=1)
        assert st == 0
        return num, den
    return compute


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    num, den = recross_sharded(_compute_factory(), NPAIRS, EVOL, rank, world)
    if rank == 0:
        torch.save((num.clone(), den), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    num2, den2 = torch.load(out)
    num1, den1 = recross_sharded(_compute_factory(), NPAIRS, EVOL, 0, 1)
    assert abs(den1 - den2) < 1e-12 * abs(den1)
    assert (num1 - num2).abs().max().item() < 1e-12 * max(1.0, num1.abs().max().item())


# ---- umbrella windows partitioned over two ranks (rate.umbrella_sampling(shard=...)) ------------------
def _umbrella(shard):
    from caracal_b200 import rate as R
    from tests import common as C
    from tests.oracle_handle import OracleRPMD
    name, nb = "h3", 2
    g = OracleRPMD(name, nb, C.masses(name), C.beta_calc_rate(300.0), C.dt_au(0.1))
    g.set_mechanism(C.mechanism(name))
    g.set_seed(C.SEED)
    g.set_thermostat(1, 5, 300.0)
    xi = np.array([0.95, 0.97, 0.99, 1.0, 1.01])
    struc = np.array([C.ring_polymer(name, 1, np.random.default_rng(k), 0.0)[0] for k in range(5)])
    return R.umbrella_sampling(g, xi, struc, 15.0, 2, 6, 12, shard=shard)


def _worker_umb(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    avg, var, _ = _umbrella((rank, world))
    if rank == 1:                                   # every rank holds the gathered statistics
        torch.save((avg, var), out)
    dist.barrier()
    dist.destroy_process_group()


def test_umbrella_windows_sharded_over_two_ranks(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "umb.pt")
    mp.spawn(_worker_umb, args=(2, port, out), nprocs=2, join=True)
    avg2, var2 = torch.load(out, weights_only=False)
    avg1, var1, _ = _umbrella(None)
    assert np.abs(avg1 - avg2).max() < 1e-13 and np.abs(var1 - var2).max() < 1e-13
    assert umbrella_sharded(lambda w0, c: (np.arange(w0, w0 + c), np.ones(c)), 4, 0, 1)[0].tolist() == [0, 1, 2, 3]


# ---- rate.recrossing(shard=...): every rank runs its block of the pair range (shard_range returns
# (start, count), not (lo, hi)) ---------------------------------------------------------------------------------
def _recross(shard):
    from caracal_b200 import rate as R
    from tests import common as C
    from tests.oracle_handle import OracleRPMD
    name, nb = "h3", 2
    g = OracleRPMD(name, nb, C.masses(name), C.beta_calc_rate(300.0), C.dt_au(0.1))
    g.set_mechanism(C.mechanism(name))
    g.set_seed(C.SEED)
    q0 = C.ring_polymer(name, nb, np.random.default_rng(3), 0.0)
    num, den, parents, st = R.recrossing(g, q0, 0.99, 0.0, 300.0, recr_equi=9, child_tot=12, child_interv=4,
                                         child_point=4, child_evol=10, shard=shard)
    return num, den


def _worker_rec(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    num, den = _recross((rank, world))
    t = torch.cat([torch.as_tensor(num, dtype=torch.float64), torch.tensor([den], dtype=torch.float64)])
    dist.all_reduce(t)
    if rank == 0:
        torch.save(t, out)
    dist.barrier()
    dist.destroy_process_group()


def test_rate_recrossing_sharded_over_three_ranks(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "rec.pt")
    mp.spawn(_worker_rec, args=(3, port, out), nprocs=3, join=True)
    t = torch.load(out)
    num1, den1 = _recross(None)
    assert den1 > 0 and abs(den1 - t[-1].item()) < 1e-12 * abs(den1)
    assert np.abs(num1 - t[:-1].numpy()).max() < 1e-12 * max(1.0, np.abs(num1).max())


def test_retry_streams_are_keyed_by_the_global_window():
    """a re-run of global window w uses the same RNG stream ids whether the windows are sharded or not"""
    from caracal_b200 import rate as R

    class Fake:
        nbeads = 1

        def __init__(self, bad_global, win0):
            self.calls, self.bad, self.win0 = [], bad_global, win0

        def umbrella_windows(self, q0, xi, kf, ntraj, equi, samp, traj_id0=0, constrain=0):
            self.calls.append((int(traj_id0), len(xi)))
            nw = len(xi)
            var = np.full((nw, ntraj), 1e-4)
            if len(self.calls) == 1 and self.win0 <= self.bad < self.win0 + nw:
                var[self.bad - self.win0, 1] = 1.0          # one trajectory of one window fails on the first pass
            return np.zeros((nw, ntraj)), var, np.zeros((nw, ntraj), dtype=np.int32)
    xi, struc = np.linspace(0.9, 1.0, 6), np.zeros((6, 3, 3))
    full = Fake(4, 0)
    R.umbrella_sampling(full, xi, struc, 1.0, 3, 1, 1, traj_id0=1000)
    part = Fake(4, 3)
    R.umbrella_sampling(part, xi[3:], struc[3:], 1.0, 3, 1, 1, traj_id0=1000, win0=3, nwin_global=6)
    assert full.calls == [(1000, 6), (1000 + (6 + 4) * 3, 1)]
    assert part.calls == [(1000 + 3 * 3, 3), (1000 + (6 + 4) * 3, 1)]
