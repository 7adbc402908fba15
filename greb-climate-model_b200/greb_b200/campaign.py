"""Large sharded ensembles (BASELINE.json configs[3], SURVEY.md 8d "config 4" / 8e).

The reference runs one `./greb <namelist>` process per member (src/greb.f90:1030-1068); a 65,536-
member perturbed-physics ensemble is 65,536 such processes, each with its own 3-year
`qflux_correction` spin-up (src/greb.f90:221, 311-364) followed by the scenario loop (:226-234).
Here rank r of `world` owns a contiguous block of members (sharding.shard_range) and walks it in
BATCHES: a perturbed member carries 40.4 MB of flux corrections on the device, so about 4,000 fit
one B200 and a rank's 8,192 members are processed as a few handles one after the other — spin-up,
scenario, results out, next batch.  Nothing is exchanged between members or ranks on the data path;
one all-reduce at the end sums the per-year ensemble moments of the annual global-mean Tsurf.

Output policy: every member's 12 x 5 monthly-mean fields are produced on the device
each simulated year; full fields come back to the host only for members whose GLOBAL index is a
multiple of `out_stride`; every member contributes its annual global means.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

from . import lib as _lib
from . import sharding

YD, XD = _lib.YD, _lib.XD


def perturbed_member(g: int):
    """member g draws its CO2 level and six physics parameters from
    numpy.random.default_rng(1000 + g).  Returns (physics_par, constant co2_ppm)."""
    rng = np.random.default_rng(1000 + g)
    p = _lib.default_physics()
    co2 = float(rng.uniform(280.0, 1120.0))
    p.kappa = float(rng.uniform(6e5, 1e6))
    p.ct_sens *= float(rng.uniform(0.8, 1.2))
    p.ce *= float(rng.uniform(0.8, 1.2))
    p.co_turb *= float(rng.uniform(0.8, 1.2))
    p.a_cloud += float(rng.uniform(-0.05, 0.05))
    p.da_ice += float(rng.uniform(-0.05, 0.05))
    return p, co2


# device bytes of one perturbed member: its own flux corrections (src/greb.f90:110, 3 x 730 fields), state,
# accumulators and the double-buffered year of monthly means
BYTES_PER_MEMBER = (3 * 730 + 5 + 6 + 2 * 12 * 5) * 96 * 48 * 4
SHARED_BYTES = 730 * 10 * 96 * 48 * 4          # forcing + spin-up targets, once per handle


def auto_batch(free_bytes: int, reserve: int = 4 << 30, multiple: int = 148) -> int:
    """largest batch (a multiple of the SM count: whole waves of one CTA per SM) whose device arrays
    fit into `free_bytes` with `reserve` left over; at least one member"""
    n = (int(free_bytes) - reserve - SHARED_BYTES) // BYTES_PER_MEMBER
    if n >= multiple:
        n -= n % multiple
    return max(1, int(n))


def plan_batches(n_local: int, batch: int) -> List[Tuple[int, int]]:
    """[start, stop) slices of a rank's members, sizes as equal as possible and <= batch."""
    if n_local < 0 or batch < 1:
        raise ValueError("plan_batches: bad arguments")
    if n_local == 0:
        return []
    nb = -(-n_local // batch)
    return [sharding.shard_range(n_local, nb, b) for b in range(nb)]


def run_sharded(total_members: int, member_fn: Callable[[int], Tuple["_lib.Physics", Sequence[float]]], forcing,
                time_flux: int, time_scnr: int, *, rank: int = 0, world: int = 1, device: int = 0,
                batch: int | None = 2048, arith: str = "exact", out_stride: int = 1024, year0: int = 1940,
                ensemble_cls=None, reduce: bool = True) -> Dict:
    """Run members [0, total_members) of an ensemble, this process doing rank `rank`'s share.

    member_fn(g) -> (physics_par, co2_ppm path) for GLOBAL member index g (what member g's namelist
    would hold).  Returns a dict with this rank's per-member annual means, the kept monthly fields,
    the device-timed kernel milliseconds and the all-reduced ensemble moments per year."""
    cls = ensemble_cls or _lib.Ensemble
    first, last = sharding.shard_range(total_members, world, rank)
    n_local = last - first
    gmean = np.zeros((n_local, time_scnr), dtype=np.float32)
    gcos = np.zeros((n_local, time_scnr), dtype=np.float32)
    flags = np.zeros(n_local, dtype=np.int32)
    kept: Dict[int, np.ndarray] = {}
    ms_spin = ms_scen = 0.0
    launches = 0
    t_setup = t_spin = t_scen = 0.0
    if not batch:                                       # size the batches from the free device memory
        import torch
        batch = auto_batch(torch.cuda.mem_get_info(device)[0])
    batches = plan_batches(n_local, batch)
    for b0, b1 in batches:
        nb = b1 - b0
        t0 = time.perf_counter()
        ens = cls(nb, device=device)
        try:
            ens.set_arithmetic(arith)
            ens.set_forcing(forcing)
            for m in range(nb):
                p, co2 = member_fn(first + b0 + m)
                ens.set_member(m, p, _lib.pad_co2(co2, max(time_scnr, 1)), year0=year0)
            ens.init()                                                   # f:176-216
            t1 = time.perf_counter()
            ens.spinup(time_flux)                                        # f:221
            ms, nl = ens.last_kernel_ms()
            ms_spin += ms
            launches += nl
            ens.reset_scenario()                                         # f:226-227
            t2 = time.perf_counter()
            keep_local = [m for m in range(nb) if (first + b0 + m) % out_stride == 0] if out_stride > 0 else []
            if time_scnr > 0:
                out, gm, gc = ens.run(time_scnr, want_output=bool(keep_local), out_members=keep_local or None)
                if not keep_local:
                    out = None
                ms, nl = ens.last_kernel_ms()
                ms_scen += ms
                launches += nl
                gmean[b0:b1] = gm
                gcos[b0:b1] = gc
                for i, m in enumerate(keep_local):
                    kept[first + b0 + m] = out[i]
            flags[b0:b1] = ens.flags()
            t3 = time.perf_counter()
        finally:
            ens.close()
        t_setup += t1 - t0
        t_spin += t2 - t1
        t_scen += t3 - t2

    # per-year [count, sum, sumsq] of the cos-lat annual global-mean Tsurf over this rank's members
    mom = np.zeros((max(time_scnr, 1), 3), dtype=np.float64)
    for y in range(time_scnr):
        mom[y] = sharding.local_moments(gcos[:, y])
    reduced = mom
    if reduce:
        reduced = _allreduce(mom, device)
    return {"first": first, "last": last, "batches": batches, "gmean": gmean, "gmean_coslat": gcos, "monthly": kept,
            "flags": flags, "kernel_ms_spinup": ms_spin, "kernel_ms_scenario": ms_scen, "launches": launches,
            "host_s": {"setup": t_setup, "spinup": t_spin, "scenario": t_scen},
            "moments": np.asarray(reduced, dtype=np.float64)}


def _allreduce(mom: np.ndarray, device: int) -> np.ndarray:
    """one all-reduce (sum) of the [years][3] moment table: NCCL on the GPU, gloo on CPU"""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return mom
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return mom
    t = torch.as_tensor(mom.copy())
    if dist.get_backend() == "nccl":
        t = t.to(f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def ensemble_mean_std(moments: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """per-year ensemble mean and standard deviation from the reduced moment table"""
    ms = [sharding.ensemble_mean_std(row) for row in np.atleast_2d(moments)]
    return np.array([m for m, _ in ms]), np.array([s for _, s in ms])
