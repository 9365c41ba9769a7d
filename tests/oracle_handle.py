"""The oracle behind the product's host interface (caracal_b200.api.RPMD): lets the host-side rate
pipeline (caracal_b200/rate.py) run on the CPU restatement, so that `-m "not gpu"` tests cover the host
logic and the GPU tests can compare a whole calc_rate run, product against oracle, stream for stream.
TEST INFRASTRUCTURE ONLY."""
import numpy as np

from oracle import oracle as O


class OracleRPMD:
    def __init__(self, pes, nbeads, mass, beta, dt):
        self.pes, self.nbeads, self.mass, self.beta, self.dt = pes, int(nbeads), np.asarray(mass, float), beta, dt
        self.natoms = len(self.mass)
        self.mech = None
        self.thermo = (0, 0, 0.0)
        self.seed = 0

    def set_mechanism(self, m):
        self.mech = m

    def set_thermostat(self, thermostat, andersen_step=0, kelvin=0.0, nose_q=0.0):
        self.thermo = (thermostat, andersen_step, kelvin)

    def set_seed(self, seed):
        self.seed = int(seed)

    def _sys(self, tid, event, k_force=None):
        s = O.System(self.pes, self.nbeads, self.mass, self.beta, self.dt)
        if self.mech is not None:
            s.set_mechanism(self.mech)
        s.set_thermostat(*self.thermo)
        s.set_rng(self.seed, int(tid), int(event))
        if k_force is not None:
            s.set_kforce(float(k_force))
        return s

    def _shape(self, q):
        return np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.nbeads, self.natoms, 3)

    def mdinit(self, q, bias_mode=0, xi_ideal=None, k_force=None, traj_id=None, event=None):
        q = self._shape(q)
        nt = q.shape[0]
        p, g, dxi = np.zeros_like(q), np.zeros_like(q), np.zeros((nt, self.natoms, 3))
        xi = np.broadcast_to(0.0 if xi_ideal is None else xi_ideal, (nt,))
        kf = np.broadcast_to(0.0 if k_force is None else k_force, (nt,))
        ev = np.zeros(nt, dtype=np.uint32) if event is None else np.array(event, dtype=np.uint32)
        for t in range(nt):
            s = self._sys(t if traj_id is None else traj_id[t], ev[t], kf[t])
            s.q[:] = q[t]
            s.mdinit(float(xi[t]), bias_mode)
            p[t], g[t], dxi[t], ev[t] = s.p, s.derivs, s.dxi, s.event()
        return p, g, dxi, ev

    def verlet(self, q, p, derivs, nsteps=1, istep0=0, constrain=-1, xi_ideal=None, k_force=None, dxi=None,
               status=None, traj_id=None, event=None):
        nt = q.size // (self.nbeads * self.natoms * 3)
        Q, P, D = (a.reshape(nt, self.nbeads, self.natoms, 3) for a in (q, p, derivs))
        xi = np.broadcast_to(0.0 if xi_ideal is None else xi_ideal, (nt,))
        kf = np.broadcast_to(0.0 if k_force is None else k_force, (nt,))
        epot, xr, st = np.zeros(nt), np.zeros(nt), np.zeros(nt, dtype=np.int32)
        for t in range(nt):
            s = self._sys(t if traj_id is None else traj_id[t], 0 if event is None else event[t], kf[t])
            s.q[:], s.p[:], s.derivs[:] = Q[t], P[t], D[t]
            if dxi is not None:
                s.dxi[:] = dxi.reshape(nt, self.natoms, 3)[t]
            for i in range(1, nsteps + 1):
                epot[t], xr[t], code = s.verlet(istep0 + i, float(xi[t]), constrain)
                st[t] |= code
                if code:
                    break
            Q[t], P[t], D[t] = s.q, s.p, s.derivs
            if dxi is not None:
                dxi.reshape(nt, self.natoms, 3)[t] = s.dxi
            if event is not None:
                event[t] = s.event()
        return epot, xr, st

    def umbrella_windows(self, q0, xi0, k_force, ntraj, equi_steps, sample_steps, traj_id0=0, constrain=0):
        q0 = self._shape(q0)
        nwin = q0.shape[0]
        avg, var, st = np.zeros((nwin, ntraj)), np.zeros((nwin, ntraj)), np.zeros((nwin, ntraj), dtype=np.int32)
        for w in range(nwin):
            for t in range(ntraj):
                s = self._sys(traj_id0 + w * ntraj + t, 0, k_force[w])
                s.q[:] = q0[w]
                s.mdinit(float(xi0[w]), 2)
                for i in range(1, equi_steps + 1):
                    st[w, t] |= s.verlet(i, float(xi0[w]), constrain)[2]
                s.gradient_all()
                xs = np.zeros(sample_steps)
                for i in range(1, sample_steps + 1):
                    _, xs[i - 1], code = s.verlet(i, float(xi0[w]), constrain)
                    st[w, t] |= code
                a = xs.sum() / sample_steps
                avg[w, t], var[w, t] = a, (xs * xs).sum() / sample_steps - a * a
        return avg, var, st

    def umbrella_window(self, q0, xi0, k_force, ntraj, equi_steps, sample_steps, traj_id0=0):
        a, v, s = self.umbrella_windows(np.asarray(q0)[None], [xi0], [k_force], ntraj, equi_steps, sample_steps, traj_id0)
        return a[0], v[0], s[0]

    def recross_children(self, q_parents, npairs, child_evol, xi_ideal, pair0=0):
        s = self._sys(0, 0)
        num, den, code = s.recross_children(q_parents, pair0, npairs, child_evol, xi_ideal, self.seed, nthreads=4)
        return num, den, np.full(npairs, code, dtype=np.int32)
