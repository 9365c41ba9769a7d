"""Multi-GPU decomposition of the RPMD work units (one process per GPU).

The reference parallelises only whole work units with an MPI master/worker scheme
(recross.f90:334-417: child +/- pairs; calc_rate.f90:1351-1376: umbrella windows) and ships
results through files or point-to-point messages.  Here every rank takes a contiguous block of
the global unit range -- units are independent, so the data path has no collective -- and the
only exchanges are the sum of the kappa(t) numerators and the denominator (child_evol+1 doubles) and
the gather of the window statistics of the umbrella phase (2 doubles per window), both all-reduced
with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
RNG streams are keyed by the GLOBAL pair index, so the reduced sums do not depend on the
number of ranks beyond floating-point summation order.
"""
import torch
import torch.distributed as dist


def comm_init_from_torch(g, device=None, group=None):
    """Gives the handle `g` (caracal_b200.RPMD) the job's NCCL communicator behind the C-ABI (crcl_comm_init): rank 0
    makes the unique id, torch.distributed ships its 128 bytes -- the role mpi_bcast plays in the Fortran drivers.
    From then on g.recross_children(_dev) / g.umbrella_windows are collective over the global unit range."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(g.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0, group=group)
    g.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    return world, rank


def shard_range(n, rank, world):
    """Contiguous block [start, start+count) of n units for this rank; blocks differ by at most 1."""
    base, rem = divmod(int(n), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def reduce_sums(t, group=None):
    """In-place sum over ranks of a tensor holding [kappa_num(0..child_evol-1), kappa_denom]."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def recross_sharded(compute, total_pairs, child_evol, rank=None, world=None, device="cpu", group=None):
    """Runs compute(pair0, npairs) -> (num[child_evol], denom) on this rank's block and reduces.

    `compute` is the rank-local work unit: RPMD.recross_children on a GPU (or, in the CPU tests,
    the oracle).  Returns (kappa_num tensor, kappa_denom float) of the whole job on every rank.
    """
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, count = shard_range(total_pairs, rank, world)
    num, den = compute(start, count)
    t = torch.zeros(child_evol + 1, dtype=torch.float64, device=device)
    t[:child_evol] = torch.as_tensor(num, dtype=torch.float64)
    t[child_evol] = float(den)
    reduce_sums(t, group)
    return t[:child_evol], float(t[child_evol])


def umbrella_sharded(compute, nwin, rank=None, world=None, device="cpu", group=None):
    """Umbrella windows partitioned over the ranks (the reference hands whole windows to MPI workers and
    collects their statistics through files, calc_rate.f90:1351-1376,1690-1734).  compute(w0, count) ->
    (average[count], variance[count]) for this rank's contiguous block of windows; every rank gets the
    (average[nwin], variance[nwin]) of all windows back: each rank fills its slice of a zero vector and
    the vectors are summed (2*nwin doubles, the only exchange of the umbrella phase)."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    start, count = shard_range(nwin, rank, world)
    t = torch.zeros(2 * nwin, dtype=torch.float64, device=device)
    if count > 0:
        avg, var = compute(start, count)
        t[start:start + count] = torch.as_tensor(avg, dtype=torch.float64)
        t[nwin + start:nwin + start + count] = torch.as_tensor(var, dtype=torch.float64)
    reduce_sums(t, group)
    return t[:nwin].cpu().numpy(), t[nwin:].cpu().numpy()
