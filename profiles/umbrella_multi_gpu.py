"""Umbrella phase of config 2 at full size over N GPUs: calc_rate CH4+H, 16 beads, 111 windows x 10
trajectories x (10 000 equilibration + 20 000 sampling) steps, the (window, trajectory) units partitioned over the ranks by
the collective form of crcl_umbrella_windows (crcl_comm_init: one NCCL all-reduce of the statistics inside the library).
Rank 0 then repeats the whole phase alone on its GPU and compares: every unit is computed by exactly one rank with an
RNG stream keyed by its global index, so the statistics must agree bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        profiles/umbrella_multi_gpu.py [out.json] [equi_steps] [sample_steps]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import caracal_b200
    from caracal_b200 import rate as R
    from caracal_b200 import systems as S
    from caracal_b200.api import beta_calc_rate, dt_au
    from caracal_b200.shard import comm_init_from_torch
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    equi, samp = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (10000, 20000)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name, nb, kelvin, ntraj = "ch4h", 16, 300.0, 10
    m, beta, dt = S.masses(name), beta_calc_rate(kelvin), dt_au(0.1)

    def handle():
        g = caracal_b200.RPMD(name, nb, m, beta, dt, device=local)
        g.set_mechanism(S.mechanism(name))
        g.set_seed(20261017)
        g.set_thermostat(1, 80, kelvin)
        return g
    g = handle()
    if world > 1:
        comm_init_from_torch(g, device=dev)
    _, _, n_all, xi = R.window_grid(-0.05, 1.05, 0.01)          # 110 windows of the key file's grid + the TS one -> 111
    xi = np.append(xi, 1.05)
    nwin = len(xi)
    kf = np.full(nwin, 0.05 * kelvin)
    ts = S.SYSTEMS[name]["ts"]()
    q0 = np.repeat(ts[None, None], nwin, axis=0).repeat(nb, axis=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    g.umbrella_windows(q0[:8], xi[:8], kf[:8], ntraj, 50, 50)   # warm-up
    barrier()
    t0 = time.perf_counter()
    avg, var, st = g.umbrella_windows(q0, xi, kf, ntraj, equi, samp, traj_id0=1 << 20)
    barrier()
    sec = time.perf_counter() - t0
    t = torch.tensor([sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    res = None
    if rank == 0:
        ref = handle()
        ref.umbrella_windows(q0[:8], xi[:8], kf[:8], ntraj, 50, 50)
        t0 = time.perf_counter()
        a1, v1, s1 = ref.umbrella_windows(q0, xi, kf, ntraj, equi, samp, traj_id0=1 << 20)
        sec1 = time.perf_counter() - t0
        bead_steps = nwin * ntraj * nb * (equi + samp)
        res = dict(n_gpus=world, windows=nwin, trajectories_per_window=ntraj, nbeads=nb, equi_steps=equi, sample_steps=samp,
                   seconds=sec, bead_steps_per_s=bead_steps / sec, seconds_one_gpu_same_box=sec1,
                   bead_steps_per_s_one_gpu=bead_steps / sec1, speedup=sec1 / sec,
                   identical_to_one_gpu=bool((avg == a1).all() and (var == v1).all() and (st == s1).all()),
                   failed_trajectories=int((st != 0).sum()), mean_abs_xi_offset=float(np.abs(avg.mean(axis=1) - xi).mean()))
        print(json.dumps(res), flush=True)
        if out_path:
            json.dump(res, open(out_path, "w"), indent=1)
    if world > 1:
        g.comm_destroy()
        dist.barrier()
        dist.destroy_process_group()
    if res is not None and not res["identical_to_one_gpu"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
