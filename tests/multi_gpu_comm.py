"""Multi-GPU check of the collective behind the C-ABI (crcl_comm_init): run under torchrun, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_comm.py

Every rank calls the work units with the GLOBAL unit range; the result on every rank must equal what one handle
without a communicator computes for the whole range (kappa sums to summation-order rounding, window statistics
bit for bit: each trajectory is computed by exactly one rank).  Prints one line per rank and exits non-zero on failure.
tests/test_gpu_round2.py::test_multi_rank_communicator launches it when the box has more than one GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import caracal_b200
    from caracal_b200 import systems as S
    from caracal_b200.api import beta_calc_rate, dt_au
    from caracal_b200.shard import comm_init_from_torch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    name, nb, seed = "ch4h", 16, 20261017

    def handle():
        g = caracal_b200.RPMD(name, nb, S.masses(name), beta_calc_rate(300.0), dt_au(0.1), device=local)
        g.set_mechanism(S.mechanism(name))
        g.set_seed(seed)
        g.set_thermostat(1, 7, 300.0)
        return g
    g, ref = handle(), handle()
    comm_init_from_torch(g, device=dev)
    assert g.comm_info()[:2] == (world, rank)
    qp = np.array([S.ring_polymer(name, nb, np.random.default_rng(k), 0.01) for k in range(3)])
    npairs, evol = 37, 60                                   # ragged: blocks differ by one pair
    num, den, st = g.recross_children(qp, npairs, evol, 0.98, pair0=11)
    num0, den0, st0 = ref.recross_children(qp, npairs, evol, 0.98, pair0=11)
    assert abs(den - den0) < 1e-12 * abs(den0), (den, den0)
    assert np.abs(num - num0).max() < 1e-12 * max(1.0, np.abs(num0).max())
    assert (st == st0).all()
    # device-pointer form, outputs resident
    d_qp = torch.as_tensor(qp, device=dev)
    d_s = torch.zeros(evol + 1, dtype=torch.float64, device=dev)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.recross_children_dev(d_qp.data_ptr(), 3, npairs, evol, 0.98, d_s.data_ptr(), d_s.data_ptr() + 8 * evol, pair0=11)
    torch.cuda.synchronize()
    s = d_s.cpu().numpy()
    assert abs(s[evol] - den0) < 1e-12 * abs(den0) and np.abs(s[:evol] - num0).max() < 1e-12 * max(1.0, np.abs(num0).max())
    # fewer pairs than ranks: some ranks run nothing and still take part in the reduction
    num1, den1, _ = g.recross_children(qp, 1, 10, 0.98, pair0=3)
    num2, den2, _ = ref.recross_children(qp, 1, 10, 0.98, pair0=3)
    assert den1 == den2 and (num1 == num2).all()
    # umbrella windows: (window, trajectory) units partitioned over the ranks
    xi = np.array([0.90, 0.95, 1.0])
    q0 = np.array([S.ring_polymer(name, nb, np.random.default_rng(k), 0.0) for k in range(3)])
    a, v, su = g.umbrella_windows(q0, xi, np.full(3, 15.0), 5, 10, 20, traj_id0=400)
    a0, v0, su0 = ref.umbrella_windows(q0, xi, np.full(3, 15.0), 5, 10, 20, traj_id0=400)
    assert (a == a0).all() and (v == v0).all() and (su == su0).all()
    g.comm_destroy()
    print("rank %d/%d ok: kappa(%d steps) = %.6f, NCCL %d" % (rank, world, evol, num[-1] / den, g.comm_info()[2]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
