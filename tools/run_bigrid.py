#!/usr/bin/env python
"""Config 5 of BASELINE.json: one member on a 0.25-degree (1440x720) grid, latitude bands over the
GPUs of a node, halo rows exchanged between neighbours (NCCL point-to-point over NVLink).

  python tools/run_bigrid.py --steps 1                                            # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29512 tools/run_bigrid.py --steps 1 --period 16

A 12-hour step here = the two circulations of the reference step (air temperature with wz_air,
humidity with wz_vapor; src/greb.f90:299-304), each nint(43200/dt_crcl) = 5,400 sub-steps at
0.25 degrees under the declared rules R1/R2 (include/greb_grid.h).  It is a strong-scaling
experiment, not a parity target: the column physics needs no neighbour data and is
not part of it.  Check before timing: N-GPU result == 1-GPU result bit for bit after --verify
sub-steps (parity with the CPU restatement, also at 1440x720, is tests/test_grid_path.py's job).
Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))


def whole_steps(args, f, rank, world, local):
    """BASELINE.json configs[4] with the column physics: N whole 12-hour steps of one member on nx x ny"""
    import torch
    import torch.distributed as dist
    import greb_b200
    from greb_b200 import bigrid
    nx, ny, n = args.nx, args.ny, args.whole_steps
    static, step_forcing_full = bigrid.s0_static_and_forcing(f, nx, ny)
    p = greb_b200.default_physics()
    f32 = np.float32
    up = (lambda a: np.ascontiguousarray(a, dtype=f32)) if (nx, ny) == (96, 48) else (lambda a: bigrid.upsample(a, ny, nx))
    toclim = np.minimum.reduce(f.tclim, axis=0)
    toclim = np.where(toclim - f32(273.15) < f32(-1.7), f32(-1.7) + f32(273.15), toclim).astype(f32)     # f:1087-1094
    cap_land = f32(f32(p.cp_land) * f32(p.rho_land)) * f32(p.d_land)
    cap_ocean = f32(p.cp_ocean) * f32(p.rho_ocean)
    z = static["z_topo"]
    state0 = {"Ts": up(f.tclim[729]), "Ta": up(f.tclim[729]), "To": up(toclim), "q": up(f.qclim[729]),
              "cap_surf": np.where(z > 0, cap_land, cap_ocean * up(f.mldclim[0])).astype(f32)}     # f:190-197

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(rk, wd, steps, timed_from):
        step_forcing = step_forcing_full
        if wd > 1 and (nx, ny) != (96, 48):          # a rank prepares only its band (+ the winds' 2 halo rows)
            lo, hi = bigrid.band_range(ny, wd, rk)
            _, step_forcing = bigrid.s0_static_and_forcing(f, nx, ny, rows=(lo - 2, hi + 2))
        big = bigrid.BigStep(nx, ny, static, state0, rank=rk, world=wd, device=local)
        barrier() if wd > 1 else torch.cuda.synchronize()
        big.run(1, timed_from - 1, step_forcing, 680.0, prefetch=False)      # warm-up
        barrier() if wd > 1 else torch.cuda.synchronize()
        t0 = time.perf_counter()
        # the forcing of step it+1 is prepared and uploaded by a worker thread while step it runs (BigStep.run)
        big.run(timed_from, steps - timed_from + 1, step_forcing, 680.0, prefetch=not args.no_prefetch)
        barrier() if wd > 1 else torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        out = {nme: big.field(nme) for nme in ("Ts", "Ta", "To", "q")}
        kms = big.kernel_ms
        if wd > 1:
            dist.barrier()
        big.close()
        return out, wall, kms

    mine, wall, kms = run(rank, world, 1 + n, 2)                 # step 1 is the warm-up
    checks = {}
    if world > 1:
        sizes = [bigrid.band_range(ny, world, r) for r in range(world)]
        pad = max(hi - lo for lo, hi in sizes)
        full = {}
        for nme, a in mine.items():
            buf = torch.zeros((pad, nx), dtype=torch.float32, device=f"cuda:{local}")
            buf[:a.shape[0]] = torch.from_numpy(a).to(f"cuda:{local}")
            allb = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(allb, buf)
            full[nme] = torch.cat([allb[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)]).cpu().numpy()
        if rank == 0:
            one, _, _ = run(0, 1, 1 + n, 2)
            checks[f"{world}_gpu_equals_1_gpu_after_{1 + n}_whole_steps"] = bool(all(np.array_equal(one[k], full[k]) for k in one))
        dist.barrier()
    t = torch.tensor([wall], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"metric": "whole 12-hour steps/s (column physics + both circulations of one member)", "value": n / float(t[0]),
                "unit": "steps/s", "n_gpus": world, "scaling": "strong", "grid": f"{nx}x{ny}", "steps_timed": n,
                "path": "column physics on tiles of 4,608 cells (greb_b200_tile_phase, the member kernel's device functions) + "
                        "persistent dataflow circulations; per step the host prepares and uploads the step's forcing ("
                        + ("in a worker thread, one step ahead" if not args.no_prefetch else "on the critical path") +
                        ") and exchanges the fields' 2 halo rows once per circulation",
                "circulation_kernel_s_per_step_rank0": kms / 1e3 / (1 + n), "mean_Ts_K": float(mine["Ts"].mean()),
                "finite": bool(all(np.isfinite(a).all() for a in mine.values())), "checks": checks}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1440)
    ap.add_argument("--ny", type=int, default=720)
    ap.add_argument("--steps", type=float, default=1.0, help="12-hour steps to time (fractions allowed)")
    ap.add_argument("--period", dest="s", type=int, default=16, help="sub-steps between halo exchanges (halo = 2*s rows)")
    ap.add_argument("--verify", type=int, default=48, help="sub-steps of the N-GPU == 1-GPU check (0 = skip)")
    ap.add_argument("--no-overlap", action="store_true", help="advance the two fields one after the other")
    ap.add_argument("--persistent", action="store_true",
                    help="v2 path: one cooperative launch per call, both fields in it, halo rows pushed GPU to GPU by "
                         "the kernel (CUDA IPC peer memory + flags); the host is not in the loop (--period is ignored)")
    ap.add_argument("--chunk", type=int, default=1350, help="--persistent: sub-steps per cooperative launch")
    ap.add_argument("--whole-steps", type=int, default=0,
                    help="time N WHOLE 12-hour steps instead: column physics on tiles (greb_b200_tile_phase) + the two "
                         "circulations on the persistent path; checks N GPUs == 1 GPU bit for bit first")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="--whole-steps: prepare every step's forcing on the critical path (the first version)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from greb_b200 import bigrid, synth
    upsample = bigrid.upsample

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("run_bigrid.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nx, ny, s = args.nx, args.ny, (1 if args.persistent else args.s)
    f = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))
    if args.whole_steps > 0:
        whole_steps(args, f, rank, world, local)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    ityr = 200
    topo = upsample(f.z_topo, ny, nx)
    fld = {"Ta": (upsample(f.tclim[ityr - 1], ny, nx), np.exp(-topo / np.float32(8400.0)).astype(np.float32)),
           "q": (upsample(f.qclim[ityr - 1], ny, nx), np.exp(-topo / np.float32(5000.0)).astype(np.float32))}
    u, v = upsample(f.uclim[ityr - 1], ny, nx), upsample(f.vclim[ityr - 1], ny, nx)
    k0, k1 = bigrid.band_range(ny, world, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(name, lo, hi):
        b = bigrid.DeviceBand(nx, ny, lo, hi, s, device=local)
        b.set_fields(fld[name][0], fld[name][1], u, v)
        return b

    checks = {}
    # ---- correctness before timing -----------------------------------------------------------
    if args.verify > 0:
        b = make("Ta", k0, k1)
        if args.persistent:
            grp = bigrid.PersistentGroup([b], rank, world)
            barrier()                                          # every rank has set its fields (and cleared its flags)
            grp.advance(args.verify)
        else:
            bigrid.advance(b, args.verify, rank, world)
        mine = torch.from_numpy(b.get()).to(f"cuda:{local}")
        barrier()                                              # nobody unmaps buffers a neighbour still writes
        b.close()
        if world > 1:                                          # gather the bands on every rank (padded: ragged heights)
            sizes = [bigrid.band_range(ny, world, r) for r in range(world)]
            pad = max(hi - lo for lo, hi in sizes)
            buf = torch.zeros((pad, nx), dtype=torch.float32, device=f"cuda:{local}")
            buf[:mine.shape[0]] = mine
            allb = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(allb, buf)
            full = torch.cat([allb[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)]).cpu().numpy()
        else:
            full = mine.cpu().numpy()
        if rank == 0:
            one = make("Ta", 0, ny)
            bigrid.advance(one, args.verify)
            checks[f"{world}_gpu_equals_1_gpu_after_{args.verify}_substeps"] = bool(np.array_equal(one.get(), full))
            one.close()

    # ---- timing ---------------------------------------------------------------------------------
    bands = {n: make(n, k0, k1) for n in ("Ta", "q")}
    nsub = bands["Ta"].nsub
    n_time = max(s, int(round(args.steps * nsub)))
    grp = None
    if args.persistent:
        grp = bigrid.PersistentGroup(list(bands.values()), rank, world)
        barrier()
        grp.advance(48)                                         # warm-up
        grp.kernel_ms, grp.launches = 0.0, 0
        n_time = max(1, int(round(args.steps * nsub)))
    else:
        for b in bands.values():                                # warm-up: 3 exchange periods
            bigrid.advance(b, 3 * s, rank, world)
            b.kernel_ms, b.launches = 0.0, 0
    barrier()
    t0 = time.perf_counter()
    if grp is not None:
        exchanges, done = 0, 0
        while done < n_time:                                    # a few cooperative launches, no host work between sub-steps
            m = min(args.chunk, n_time - done)
            grp.advance(m)
            done += m
    elif args.no_overlap:
        exchanges = sum(bigrid.advance(b, n_time, rank, world) for b in bands.values())
    else:                                                       # one field's exchange behind the other's sub-steps
        exchanges = bigrid.advance_overlapped(list(bands.values()), n_time, rank, world)
    barrier()
    wall = time.perf_counter() - t0
    kms = grp.kernel_ms if grp is not None else sum(b.kernel_ms for b in bands.values())
    launches = grp.launches if grp is not None else sum(b.launches for b in bands.values())
    mean_T = float(bands["Ta"].get().astype(np.float64).mean())
    t = torch.tensor([wall, kms], dtype=torch.float64, device=f"cuda:{local}")
    m = torch.tensor([mean_T * (k1 - k0)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(m)                                      # the global mean: the only collective
    if rank == 0:
        wall, kms = float(t[0]), float(t[1])
        steps_done = n_time / nsub
        line = {
            "metric": "12-hour steps/s (circulations of one 0.25-degree member)", "value": steps_done / wall,
            "unit": "steps/s", "n_gpus": world, "scaling": "strong",
            "config": {"workload": "configs[4]: synthetic 1440x720 grid, single member, latitude bands, "
                                   "S0 forcing bilinearly upsampled", "grid": f"{nx}x{ny}",
                       "substeps_per_circulation": nsub, "dt_crcl_s": bands["Ta"].dt_crcl,
                       "substeps_between_exchanges": s, "halo_rows": 2 * s,
                       "rules": "R1 dt_crcl = 1800*(48/ydim)^2; R2 |lat| clamped to 88.125 deg in dxlat, dtdff2 >= 1 s",
                       "exchange": ("inside the persistent kernel: the two outermost rows of each band are stored into the "
                                    "neighbour's halo rows (CUDA IPC peer memory over NVLink) + per-row release/acquire "
                                    "flags, every sub-step; all-reduce of the global mean at the end" if grp is not None else
                                    "NCCL point-to-point of 2*s rows per neighbour every s sub-steps; "
                                    "all-reduce of the global mean at the end"),
                       "path": "persistent cooperative kernel (v2)" if grp is not None else "one launch per sub-step (v1)",
                       "overlap": ("both fields in one kernel" if grp is not None else
                                   "none" if args.no_overlap else "the exchange of one field runs behind the sub-steps of the other")},
            "timed_substeps_per_field": n_time, "wall_s": wall, "kernel_s_max_rank": kms / 1e3,
            "cell_substeps_per_s": 2 * n_time * nx * ny / wall,
            "gpu_launches_rank0": launches, "halo_exchanges_rank0": exchanges,
            "halo_bytes_per_exchange_per_neighbour": 2 * s * nx * 4,
            "global_mean_Ta": float(m[0]) / ny, "checks": checks,
        }
        print(json.dumps(line), flush=True)
    barrier()
    for b in bands.values():
        b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
