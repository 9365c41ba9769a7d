"""Host side of calc_rate.x on the GPU path: the phases of calc_rate.f90 expressed through the C-ABI.

    phase 1  start structures for all umbrella windows      calc_rate.f90:651-1148   generate_start_structures
    phase 2  umbrella equilibration + sampling              calc_rate.f90:1253-1734  umbrella_sampling
    phase 3  umbrella integration (Kaestner & Thiel) -> PMF calc_rate.f90:1935-1997  umbrella_integration
             barrier / reactant bins                        calc_rate.f90:2056-2091  locate_extrema
    phase 4  recrossing: constrained parent + child pairs   recross_serial.f90:83-307 / recross.f90   recrossing
    phase 5  k(T)                                           calc_k_t.f90:101-137     calc_k_t

The reference runs phases 2 and 4 as MPI master/worker loops over whole work units (windows, child
pairs); here a work unit is a trajectory of a batch and the batch is one kernel launch
(crcl_umbrella_windows, crcl_recross_children).  Serial Markov chains (phase 1, the recrossing parent)
stay serial, as in the reference, but run inside one launch per segment.  Nothing here touches oracle/.

Deliberate differences, all statistical only (the reference's own RNG stream is not reproducible):
  * the umbr_traj trajectories of a window start from the window's start structure independently;
    the reference chains them (trajectory j starts where j-1 ended, calc_rate.f90:1383-1387 sets q_i
    only once per window);
  * failed trajectories (status != 0 or var(xi) > 1E-2, calc_rate.f90:1679) are re-run with fresh RNG
    streams instead of the rpmd_check.f90 start-structure shifting.
"""
import math

import numpy as np

HARTREE_KJ = 2625.50  # calc_rate.f90:1992 (pmf written in kJ/mol)


def window_grid(umbr_lo, umbr_hi, umbr_dist):
    """calc_rate.f90:655-665: n_over, n_samplings, n_all and the xi of windows 1..n_all-1."""
    n_over = int(round((umbr_hi - 1.0) / umbr_dist))
    n_samplings = int(round((1.0 - umbr_lo) / umbr_dist))
    n_all = int(round((umbr_hi - umbr_lo) / umbr_dist)) + 1
    if (n_over + n_samplings) - (n_all - 1) != 0:
        raise ValueError("umbrella step size does not fit into the umbrella bounds (calc_rate.f90:660)")
    xi = np.zeros(n_all - 1)
    for i in range(1, n_over + 1):
        xi[n_all + i - (n_over + 1) - 1] = 1.0 + (i - 1) * umbr_dist
    for i in range(1, n_samplings + 1):
        xi[n_samplings - i] = 1.0 - i * umbr_dist
    return n_over, n_samplings, n_all, xi


def generate_start_structures(g1, ts_xyz, mass, xi_wins, n_over, n_samplings, k_force, gen_steps, traj_id0=0,
                              constrain=0):
    """Phase 1 with a ONE-bead handle g1 (calc_rate.f90:651 sets nbeads = 1): from the TS structure up
    through the windows xi >= 1, then from the last xi = 1 structure down to umbr_lo; every window runs
    mdinit(bias) + gen_steps biased steps and hands its last structure to the next one.  Returns
    struc_equi[nwin, natoms, 3] (centre of mass removed, :1118-1130) and the xi reached."""
    nwin = len(xi_wins)
    natoms = len(mass)
    struc = np.zeros((nwin, natoms, 3))
    start_xis = np.zeros(nwin)
    k_force = np.broadcast_to(np.asarray(k_force, dtype=np.float64), (nwin,))

    def run(q, w):
        tid = np.array([traj_id0 + w], dtype=np.uint32)
        xi0 = np.array([xi_wins[w]])
        kf = np.array([k_force[w]])
        p, d, dxi, ev = g1.mdinit(q, 2, xi_ideal=xi0, k_force=kf, traj_id=tid)
        ep, xr, st = g1.verlet(q, p, d, nsteps=gen_steps, constrain=constrain, xi_ideal=xi0, k_force=kf, dxi=dxi,
                               traj_id=tid, event=ev)
        if st[0] != 0:
            raise RuntimeError("start-structure trajectory of window %d failed (status %d)" % (w, st[0]))
        struc[w] = q[0, 0]
        start_xis[w] = xr[0]

    q = np.array(ts_xyz, dtype=np.float64).reshape(1, 1, natoms, 3).copy()
    first_over = nwin - n_over
    for w in range(first_over, nwin):
        run(q, w)
    q[0, 0] = struc[first_over] if n_over > 0 else np.asarray(ts_xyz, dtype=np.float64)
    for w in range(n_samplings - 1, -1, -1):
        run(q, w)
    m = np.asarray(mass, dtype=np.float64)
    struc -= (struc * m[None, :, None]).sum(axis=1, keepdims=True) / m.sum()
    return struc, start_xis


def umbrella_sampling(g, xi_wins, struc_equi, k_force, umbr_traj, equi_steps, umbr_steps, traj_id0=1 << 20,
                      max_retry=5, constrain=0, shard=None, device="cpu", win0=0, nwin_global=None, per_trajectory=False,
                      window_ids=None):
    """Phase 2: all windows x umbr_traj trajectories in one batch (crcl_umbrella_windows); returns the
    window averages and variances of xi as statistics/bias_* hold them (calc_rate.f90:1690-1700).
    shard = (rank, world): the windows are partitioned over the ranks (one GPU each) and the statistics
    gathered with one all-reduce.  RNG streams are keyed by the GLOBAL window index -- trajectory t of global window
    w uses stream traj_id0 + w*umbr_traj + t, and its r-th re-run traj_id0 + r*nwin_global*umbr_traj + w*umbr_traj + t
    -- so neither the first pass nor a retry depends on the number of ranks (win0: global index of xi_wins[0]).
    window_ids: the global indices of the given windows when they are a non-contiguous subset (restart of a run with
    some statistics files already complete): every window is then its own launch with its global stream ids.
    per_trajectory: return the [nwin, umbr_traj] arrays (one line each of statistics/bias_<xi>) instead of the means."""
    if shard is not None:
        from .shard import umbrella_sharded, reduce_sums
        import torch
        kf_all = np.broadcast_to(np.asarray(k_force, dtype=np.float64), (len(xi_wins),))
        nerr_box = [0]

        def compute(w0, cnt):
            a, v, ne = umbrella_sampling(g, xi_wins[w0:w0 + cnt], struc_equi[w0:w0 + cnt], kf_all[w0:w0 + cnt],
                                         umbr_traj, equi_steps, umbr_steps, traj_id0=traj_id0,
                                         max_retry=max_retry, constrain=constrain, win0=w0, nwin_global=len(xi_wins))
            nerr_box[0] = ne
            return a, v
        avg, var = umbrella_sharded(compute, len(xi_wins), shard[0], shard[1], device=device)
        ne = torch.tensor([float(nerr_box[0])], dtype=torch.float64, device=device)
        reduce_sums(ne)                       # the count of re-run trajectories of the whole job
        return avg, var, int(ne.item())
    nwin = len(xi_wins)
    nglob = nwin if nwin_global is None else int(nwin_global)
    k_force = np.broadcast_to(np.asarray(k_force, dtype=np.float64), (nwin,)).copy()
    q0 = np.repeat(struc_equi[:, None], g.nbeads, axis=1)
    gid = (win0 + np.arange(nwin)) if window_ids is None else np.asarray(window_ids, dtype=np.int64)
    if window_ids is None or (np.diff(gid) == 1).all():
        avg, var, st = g.umbrella_windows(q0, xi_wins, k_force, umbr_traj, equi_steps, umbr_steps,
                                          traj_id0=traj_id0 + int(gid[0]) * umbr_traj, constrain=constrain)
    else:
        avg, var = np.zeros((nwin, umbr_traj)), np.zeros((nwin, umbr_traj))
        st = np.zeros((nwin, umbr_traj), dtype=np.int32)
        for w in range(nwin):
            avg[w:w + 1], var[w:w + 1], st[w:w + 1] = g.umbrella_windows(
                q0[w:w + 1], xi_wins[w:w + 1], k_force[w:w + 1], umbr_traj, equi_steps, umbr_steps, constrain=constrain,
                traj_id0=traj_id0 + int(gid[w]) * umbr_traj)
    nerr = 0
    for r in range(max_retry):
        bad = (st != 0) | ~(var <= 1e-2)      # calc_rate.f90:1679 (1E-2 is a REAL*4 literal; var is far away)
        wbad = np.nonzero(bad.any(axis=1))[0]
        if len(wbad) == 0:
            break
        nerr += int(bad.sum())
        for w in wbad:                        # rare: one call per window keeps the stream ids global
            a2, v2, s2 = g.umbrella_windows(q0[w:w + 1], xi_wins[w:w + 1], k_force[w:w + 1], umbr_traj, equi_steps,
                                            umbr_steps, constrain=constrain,
                                            traj_id0=traj_id0 + ((r + 1) * nglob + int(gid[w])) * umbr_traj)
            sel = bad[w]
            avg[w, sel], var[w, sel], st[w, sel] = a2[0, sel], v2[0, sel], s2[0, sel]
    else:
        if ((st != 0) | ~(var <= 1e-2)).any():
            raise RuntimeError("umbrella trajectories keep failing (rpmd_check.f90 would call fatal)")
    if per_trajectory:
        return avg, var, nerr
    return avg.mean(axis=1), var.mean(axis=1), nerr


def umbrella_integration(xi_wins, average, variance, k_force, beta, xi_min, xi_max, nbins, umbr_traj, umbr_step):
    """Phase 3, calc_rate.f90:1935-1997: mean force of every bin as the normal-distribution weighted
    average of the window mean forces, trapezoid integration, minimum shifted to zero.  Returns
    bin_coord[nbins] and pmf[nbins] in hartree; element nbins-1 is never assigned in the reference and
    stays 0 before the shift (it takes part in minval and maxloc) -- reproduced."""
    xi_wins, average, variance = (np.asarray(a, dtype=np.float64) for a in (xi_wins, average, variance))
    k_force = np.broadcast_to(np.asarray(k_force, dtype=np.float64), xi_wins.shape)
    bin_size = (xi_max - xi_min) / (nbins - 1)
    xi_act = xi_min + bin_size * np.arange(nbins)
    dA = np.zeros(nbins)
    nw = float(umbr_traj * umbr_step)
    for j in range(nbins):
        p_ib = 1.0 / np.sqrt(2.0 * math.pi * variance) * np.exp(-0.5 * (xi_act[j] - average) ** 2 / variance)
        dA_iu = (1.0 / beta) * (xi_act[j] - average) / variance - k_force * (xi_act[j] - xi_wins)
        denom = 0.0
        for k in range(len(xi_wins)):
            denom = denom + nw * p_ib[k]
        acc = 0.0
        for k in range(len(xi_wins)):
            acc = acc + nw * p_ib[k] * dA_iu[k]
        dA[j] = acc / denom
    pmf = np.zeros(nbins)
    bin_coord = np.zeros(nbins)
    A = 0.0
    for i in range(nbins - 1):
        bin_coord[i] = 0.5 * (xi_min + i * bin_size + xi_min + (i + 1) * bin_size)
        A = A + 0.5 * bin_size * (dA[i] + dA[i + 1])
        pmf[i] = A
    pmf = pmf - pmf.min()
    return bin_coord, pmf


def locate_extrema(bin_coord, pmf, xi_min, xi_max, pmf_minloc="ZERO", xi_pos_manual=None):
    """calc_rate.f90:2056-2091.  Returns (maxlocate, minlocate, xi_barrier), 0-based bin indices;
    maxlocate is the PMF maximum even when the recrossing plane is moved by hand (:2064-2075)."""
    nbins = len(pmf)
    maxlocate = int(np.argmax(pmf))
    xi_barrier = bin_coord[maxlocate]
    rec = maxlocate
    if xi_pos_manual is not None:
        xi_barrier = xi_pos_manual
        rec = int(np.argmin(np.abs(bin_coord - xi_pos_manual)))
    if pmf_minloc == "ZERO":
        minlocate = -int(xi_min / ((xi_max - xi_min) / nbins))      # Fortran: -int(...)+1, 1-based
    else:
        minlocate = int(np.argmin(pmf[:rec + 1]))
    return maxlocate, minlocate, xi_barrier


def recrossing(g, q_start, xi_barrier, k_force, kelvin, recr_equi, child_tot, child_interv, child_point, child_evol,
               traj_id0=1 << 24, shard=None, folder=None, rounds_per_launch=None):
    """Phase 4 (recross_serial.f90:83-307): constrained parent equilibrated for recr_equi steps, then
    child_times = child_tot/child_point spawn points child_interv constrained steps apart, child_point/2
    +/- pairs each, child_evol free steps per child.  Returns kappa(t)[child_evol], num, denom.
    shard = (rank, world): this rank evaluates its contiguous block of the pair range (the reference
    hands out pairs to MPI workers, recross.f90:334-417); sums must then be added across ranks.
    folder (rate_io.RunFolder): keep the reference's restart files and resume from them."""
    child_times = child_tot // child_point
    npp = child_point // 2
    q = np.array(q_start, dtype=np.float64).reshape(1, g.nbeads, g.natoms, 3).copy()
    xi = np.array([xi_barrier])
    kf = np.array([float(k_force)])
    tid = np.array([traj_id0], dtype=np.uint32)
    # parent equilibration: Andersen every int(sqrt(recr_equi)) steps (recross_serial.f90:92-95)
    g.set_thermostat(1, int(math.sqrt(float(recr_equi))), kelvin)
    p, d, dxi, ev = g.mdinit(q, 2, xi_ideal=xi, k_force=kf, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=recr_equi, constrain=1, xi_ideal=xi, k_force=kf, dxi=dxi, traj_id=tid,
                          event=ev)
    if st[0] != 0:
        raise RuntimeError("recrossing parent failed during equilibration (status %d)" % st[0])
    parents = np.zeros((child_times, g.nbeads, g.natoms, 3))
    a_step = int(math.sqrt(float(child_interv)))
    for i in range(1, child_times + 1):
        parents[i - 1] = q[0]
        if i == child_times:
            break               # the reference still propagates the parent once more; nobody reads it
        # recross_serial.f90:283: verlet is called with istep = i for the whole segment, so Andersen
        # fires on EVERY step when mod(i, andersen_step) == 0 and never otherwise
        g.set_thermostat(1, 1 if (a_step > 0 and i % a_step == 0) else 0, kelvin)
        p, d, dxi, ev = g.mdinit(q, 2, xi_ideal=xi, k_force=kf, traj_id=tid, event=ev)
        ep, xr, st = g.verlet(q, p, d, nsteps=child_interv, constrain=1, xi_ideal=xi, k_force=kf, dxi=dxi,
                              traj_id=tid, event=ev)
        if st[0] != 0:
            raise RuntimeError("recrossing parent failed in segment %d (status %d)" % (i, st[0]))
    # children: pair g belongs to parent g mod child_times; no thermostat, no bias (:131-135)
    npairs = child_times * npp
    g.set_thermostat(0, 0, kelvin)
    if folder is not None and shard is None:
        # restartable form (recross.f90:134-226,420-440): the pair range is processed in launches of whole ROUNDS
        # (round r = the r-th pair of every parent = pairs [r*child_times, (r+1)*child_times)), so any chunking gives
        # the sums of the single launch; after every launch the accumulated sums and the bunch count go to the files
        # the reference keeps (a round is child_times/npp of its "bunches")
        done_b, num, den, _ = folder.recross_resume(child_evol, g.nbeads, g.natoms)
        r0 = (done_b * npp) // child_times if (done_b * npp) % child_times == 0 else 0
        if r0 == 0:
            num, den = np.zeros(child_evol), 0.0
        rounds = npp
        step = rounds if not rounds_per_launch else int(rounds_per_launch)
        status = np.zeros(npairs, dtype=np.int32)
        while r0 < rounds:
            r1 = min(rounds, r0 + step)
            n_, d_, st_ = g.recross_children(parents, (r1 - r0) * child_times, child_evol, xi_barrier, pair0=r0 * child_times)
            num, den = num + n_, den + d_
            status[r0 * child_times:r1 * child_times] = st_
            r0 = r1
            folder.recross_checkpoint((r0 * child_times) // npp, num, den, q[0])
        folder.write_recrossing_time(num, den, g.dt)
        folder.recross_finished(num[-1] / den)
        return num, den, parents, status
    lo, cnt = 0, npairs
    if shard is not None:
        from .shard import shard_range
        lo, cnt = shard_range(npairs, shard[0], shard[1])      # (start, count)
    num, den, status = g.recross_children(parents, cnt, child_evol, xi_barrier, pair0=lo)
    return num, den, parents, status


def calc_k_t(kappa, pmf_max, pmf_min, beta, mass_reac, R_inf, npaths):
    """calc_k_t.f90:101-137, bimolecular: k(T) in cm^3/(mol s) and cm^3/(molecule s).  mass_reac in
    atomic units (electron masses), as module evb_mod holds them."""
    my_R = mass_reac[0] * mass_reac[1] / (mass_reac[0] + mass_reac[1])
    k_t = npaths * kappa * 4.0 * math.pi * R_inf * R_inf * math.sqrt(1.0 / (2.0 * math.pi * beta * my_R)) \
        * math.exp(-beta * (pmf_max - pmf_min))
    # calc_k_t.f90:125: the unit factors are REAL*4 literals, combined in single precision (F3)
    b = np.float32(5.2917721092e-11)
    unit = np.float32(np.float32(b * b) * b) / np.float32(2.418884326505e-17)
    k_t = k_t * 1e6 * float(unit) * float(np.float32(6.02214179E23))
    avogadro = 6.02214179e23
    return k_t, k_t / avogadro


def calc_k_t_unimol(kappa, pmf_max, pmf_min, kelvin, npaths):
    """calc_k_t.f90:199-262, CYCLOREVER / REARRANGE / DECOM_1BOND / ELIMINATION: the simple TST formula
    k(T) = kappa n_paths k_B T / h exp(-(W(xi_TS) - W(xi_min)) / (k_B T)) in s^-1.  pmf_* in hartree; the
    constants are the reference's REAL*4 literals (F3)."""
    f32 = lambda x: float(np.float32(x))
    w_max, w_min = pmf_max * 2625.50, pmf_min * 2625.50          # kJ/mol (:235-236)
    return kappa * npaths * f32(1.3806485E-23) * kelvin / f32(6.62607E-34) * \
        math.exp(-(w_max - w_min) / (f32(0.00831447) * kelvin))


def calc_rate(g, g1, ts_xyz, mass, mech, kelvin, beta, umbr_lo=-0.05, umbr_hi=1.05, umbr_dist=0.01, k_force_all=0.05,
              gen_steps=10000, equi_steps=10000, umbr_steps=20000, umbr_traj=10, xi_min=-0.05, xi_max=1.05,
              nbins=5000, recr_equi=50000, child_tot=10000, child_interv=1000, child_point=100, child_evol=500,
              andersen_step=80, npaths=1, pmf_minloc="ZERO", umbr_constrain=0, log=None, workdir=None, names=None,
              rounds_per_launch=None):
    """The whole calc_rate.x run (defaults = examples/calc_rate/h+h2/rate.key).  g: handle with the
    ring-polymer bead count, g1: one-bead handle of the same system (phase 1); both need
    set_mechanism and set_seed.  umbr_constrain: the constrain flag of the biased phases 1 and 2, 0 as
    calc_rate.f90 passes it, 3 for the same dynamics without verlet.f90:1300-1306's removal of net
    rotation ("published figures").  workdir: keep the reference's files and restart markers under
    workdir/<T>K_<n>bead/ (rate_io.RunFolder) and resume from them -- finished phases are read back instead of run
    again, as calc_rate.f90 does.  Returns a dict with every intermediate."""
    import time
    t_prev = [time.perf_counter()]
    timings = {}

    def say(msg, phase=None):
        now = time.perf_counter()
        if phase:
            timings[phase] = now - t_prev[0]
            msg += "  [%.2f s]" % timings[phase]
        t_prev[0] = now
        if log:
            log(msg)
    n_over, n_samplings, n_all, xi_wins = window_grid(umbr_lo, umbr_hi, umbr_dist)
    k_force = np.full(n_all - 1, k_force_all * kelvin)          # calc_rate.f90:699
    folder = None
    if workdir is not None:
        from .rate_io import RunFolder
        folder = RunFolder(workdir, kelvin, g.nbeads)
    # ---- phase 1 (skipped when the marker of a previous run is there, calc_rate.f90:751,1153-1163)
    if folder is not None and folder.has("start_finished"):
        xi_file, struc = folder.read_start_structures()
        start_xis = xi_file.copy()
        say("start structures: read from %s" % folder.f("equilibrated_struc.xyz"), "start_structures")
    else:
        g1.set_thermostat(1, andersen_step, kelvin)
        struc, start_xis = generate_start_structures(g1, ts_xyz, mass, xi_wins, n_over, n_samplings, k_force, gen_steps,
                                                     constrain=umbr_constrain)
        if folder is not None:
            folder.write_start_structures(names or ["X"] * len(mass), xi_wins, start_xis, struc)
            _, struc = folder.read_start_structures()          # the reference always continues from the file (:1302-1313)
        say("start structures: xi reached in [%.3f, %.3f]" % (start_xis.min(), start_xis.max()), "start_structures")
    # ---- phase 2 (statistics/bias_<xi> per window; windows already on file are not run again, :1420-1478)
    g.set_thermostat(1, andersen_step, kelvin)
    if folder is None:
        average, variance, nerr = umbrella_sampling(g, xi_wins, struc, k_force, umbr_traj, equi_steps, umbr_steps,
                                                    constrain=umbr_constrain)
    else:
        nerr = 0
        if not folder.has("sampling_finished"):
            todo = [w for w in range(len(xi_wins)) if folder.stats_resume(xi_wins[w], umbr_traj)[0] <= umbr_traj]
            if todo:
                # a window that was interrupted half way is run again as a whole (its trajectories are one launch here)
                a_t, v_t, nerr = umbrella_sampling(g, xi_wins[todo], struc[todo], k_force[todo], umbr_traj, equi_steps,
                                                   umbr_steps, constrain=umbr_constrain, per_trajectory=True,
                                                   window_ids=todo, nwin_global=len(xi_wins))
                for i, w in enumerate(todo):
                    folder.stats_write(xi_wins[w], 1, a_t[i], v_t[i], umbr_traj)
            from .rate_io import touch
            touch(folder.f("sampling_finished"))
        average, variance = folder.stats_read(xi_wins, umbr_traj)
        folder.write_umbr_int(xi_wins, average, variance)
    say("umbrella sampling: %d windows, %d re-run trajectories" % (len(xi_wins), nerr), "umbrella_sampling")
    bin_coord, pmf = umbrella_integration(xi_wins, average, variance, k_force, beta, xi_min, xi_max, nbins, umbr_traj,
                                          umbr_steps)
    if folder is not None:
        folder.write_pmf(bin_coord, pmf)
    maxloc, minloc, xi_barrier = locate_extrema(bin_coord, pmf, xi_min, xi_max, pmf_minloc)
    say("PMF: barrier %.3f kJ/mol at xi = %.4f" % ((pmf[maxloc] - pmf[minloc]) * HARTREE_KJ, xi_barrier), "umbrella_integration")
    ts_locate = int(np.argmin(np.abs(xi_wins - xi_barrier)))     # calc_rate.f90:2183-2187
    q_start = np.repeat(struc[ts_locate][None], g.nbeads, axis=0)
    num, den, parents, status = recrossing(g, q_start, xi_barrier, k_force[ts_locate], kelvin, recr_equi, child_tot,
                                           child_interv, child_point, child_evol, folder=folder,
                                           rounds_per_launch=rounds_per_launch)
    kappa_t = num / den
    kappa = kappa_t[-1]
    if kappa < 0.002:                                            # calc_rate.f90:2236-2252
        kappa = 1.0
    mass_reac = [sum(mass[a - 1] for a in r) for r in mech.reactants]
    k_t, k_t_molec = calc_k_t(kappa, pmf[maxloc], pmf[minloc], beta, mass_reac, mech.R_inf, npaths)
    say("kappa = %.4f, k(T) = %.4e cm^3/(molecule s)" % (kappa, k_t_molec), "recrossing")
    return dict(xi_wins=xi_wins, struc_equi=struc, start_xis=start_xis, average=average, variance=variance,
                bin_coord=bin_coord, pmf=pmf, maxlocate=maxloc, minlocate=minloc, xi_barrier=xi_barrier,
                delta_w_kj=(pmf[maxloc] - pmf[minloc]) * HARTREE_KJ, kappa_t=kappa_t, kappa=kappa, k_t=k_t,
                k_t_molec=k_t_molec, child_status=status, n_rerun=nerr, timings=timings)
