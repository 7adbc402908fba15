#!/usr/bin/env python
"""Config 4 of BASELINE.json: a perturbed-physics ensemble sharded over the GPUs of one node.

  python tools/run_ensemble.py --members 65536 --time-flux 3 --time-scnr 50            # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29511 tools/run_ensemble.py --members 65536 --time-flux 3 --time-scnr 50

Every member does what one `./greb <namelist>` process does in the reference (src/greb.f90:
1030-1098): its own `time_flux`-year flux-correction spin-up, then `time_scnr` scenario years.
Rank 0 prints ONE JSON line: member-years/s of the scenario phase (CUDA events on the launch
stream, max over ranks), of spin-up + scenario, and of the whole job by the wall clock (host set-up,
H2D/D2H and the final NCCL all-reduce of the ensemble moments included).
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=65536, help="total members over all ranks")
    ap.add_argument("--time-flux", type=int, default=3)
    ap.add_argument("--time-scnr", type=int, default=50)
    ap.add_argument("--batch", type=int, default=2048,
                    help="members per handle (40.4 MB of corrections each); 0 = as many as fit the free device memory")
    ap.add_argument("--arith", default="exact", choices=["fast", "exact"],
                    help="exact (default): bit-identical to the reference arithmetic; fast: see DESIGN.md section 4")
    ap.add_argument("--out-stride", type=int, default=1024, help="members = 0 mod this keep their monthly fields")
    ap.add_argument("--save", default=None, help="npz file for rank 0's annual means and the ensemble statistics")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import greb_b200
    from greb_b200 import campaign, flops as fm, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("run_ensemble.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    forcing = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0 = time.perf_counter()
    r = campaign.run_sharded(args.members, campaign.perturbed_member, forcing, args.time_flux, args.time_scnr,
                             rank=rank, world=world, device=local, batch=args.batch, arith=args.arith,
                             out_stride=args.out_stride)
    barrier()
    wall = time.perf_counter() - t0

    t = torch.tensor([r["kernel_ms_scenario"], r["kernel_ms_spinup"], wall, r["host_s"]["setup"],
                      float(r["flags"].sum())], dtype=torch.float64, device=f"cuda:{local}")
    tsum = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms_scen, ms_spin, wall_max, setup_max = (float(x) for x in t[:4])
        M, Y, F = args.members, args.time_scnr, args.time_flux
        mean, std = campaign.ensemble_mean_std(r["moments"])
        p0, _ = campaign.perturbed_member(0)
        flops_my = fm.flops_per_member_year(p0.pi, 8e5)
        value = M * Y / (ms_scen / 1e3) if ms_scen > 0 else None
        peak = world * 148 * 128 * 2 * 1965e6 / 1e12
        line = {
            "metric": "member-years/sec", "value": value, "unit": "member-years/s", "n_gpus": world,
            "config": {"workload": "configs[3]: perturbed-physics ensemble sharded across the GPUs, members drawn "
                                   "from default_rng(1000+m), 96x48, synthetic S0 forcing",
                       "members": M, "members_per_gpu": M // world, "batch": args.batch, "time_flux": F, "time_scnr": Y,
                       "arithmetic": args.arith, "out_stride": args.out_stride,
                       "collective": "one NCCL all-reduce of [years][count,sum,sumsq] at the end"},
            "scenario_kernel_s": ms_scen / 1e3, "spinup_kernel_s": ms_spin / 1e3,
            "spinup_plus_scenario_member_years_per_s": M * (Y + F) / ((ms_scen + ms_spin) / 1e3),
            "wall_s": wall_max, "host_setup_s": setup_max,
            "whole_job_member_years_per_s": M * (Y + F) / wall_max,
            "roofline": {"bound": "fp32", "achieved": (value or 0) * flops_my / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": (value or 0) * flops_my / 1e12 / peak,
                         "note": "as-written flops per member-year at the default kappa x value; peak = n_gpus x 148 x "
                                 "128 x 2 x 1965 MHz"},
            "gpu_launches_rank0": r["launches"], "nonfinite_members": int(float(tsum[4])),
            "kept_monthly_members_rank0": sorted(r["monthly"].keys()),
            "ensemble_gmean_tsurf_coslat": {"first_year": [float(mean[0]), float(std[0])],
                                            "last_year": [float(mean[-1]), float(std[-1])],
                                            "count": float(r["moments"][0, 0])},
        }
        print(json.dumps(line), flush=True)
        if args.save:
            np.savez_compressed(args.save, gmean=r["gmean"], gmean_coslat=r["gmean_coslat"], moments=r["moments"],
                                mean=mean, std=std)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
