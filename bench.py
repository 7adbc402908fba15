#!/usr/bin/env python
"""bench.py — member-years/sec of the B200-native GREB stepping core (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[2], SURVEY.md 8d "config 3 (ii)"): a 1,024-member perturbed-
parameter + CO2 ensemble per GPU on the synthetic S0 forcing; member m draws its parameters from
numpy.random.default_rng(1000+m).  One bench STEP = one simulated scenario year (730 twelve-hour
steps, 12 month-end outputs of 5 fields) for every member of the rank.  Weak scaling: every GPU
gets its own 1,024 members; the only collective is the NCCL all-reduce of the ensemble sum /
sum-of-squares of the per-member annual global-mean Tsurf (a few floats per year).

`value` is timed with CUDA events on the kernels' launch stream with all inputs resident in HBM
and the monthly means written to the device ring buffer; `e2e` is the same year-step through the
C ABI with HOST buffers: H2D of every member's state from pinned memory, the run, D2H of all
monthly means and the diagnostics.  Flux corrections (41 GB for 1,024 members) exceed the 126 MB
L2 many times over, so successive timed steps cannot hit in L2 ("inputs larger than L2").
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "greb-climate-model_b200"))
sys.path.insert(0, ROOT)

MEMBERS_PER_GPU = 1024
NCU_FILES = {"exact": "profiles/r02_member_kernel_exact_ncu_summary.txt",
             "fast": "profiles/r02_member_kernel_fast_ncu_summary.txt"}


def ncu_counters(arith: str):
    """Counters of the committed `ncu --set full` capture of greb_member_kernel (tools/ncu_summary.py output:
    a header with the command, the kernel name and the members per launch, then `metric [unit] = value` lines).
    Read at run time — nothing is typed into this file; returns None when the capture is missing."""
    path = os.path.join(ROOT, NCU_FILES[arith])
    if not os.path.exists(path):
        return None
    vals, head = {}, []
    for ln in open(path):
        ln = ln.rstrip("\n")
        if " = " in ln and "[" in ln.split(" = ")[0]:
            name = ln.split(" [")[0].strip()
            unit = ln.split("[")[1].split("]")[0]
            try:
                v = float(ln.split(" = ")[1])
            except ValueError:
                continue
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
            vals[name] = v * scale
        elif not vals:
            head.append(ln)
    import re
    m = re.search(r"(\d+) members", " ".join(head))
    k = re.search(r"kernel: (\S+)", " ".join(head))
    n_members = int(m.group(1)) if m else None
    if not n_members or "smsp__inst_executed.sum" not in vals:
        return None
    out = {"source": NCU_FILES[arith], "capture": head[0] if head else "", "kernel": k.group(1) if k else None,
           "members_per_launch": n_members,
           "dram_bytes_per_member_year": (vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)) / n_members,
           "warp_instructions_per_member_year": vals["smsp__inst_executed.sum"] / n_members,
           "ipc_per_sm": vals.get("sm__inst_executed.avg.per_cycle_active"),
           "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "fma_pipe_inst_pct": vals.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
           "alu_pipe_inst_pct": vals.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
           "lsu_wavefronts_pct": vals.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
           "registers_per_thread": vals.get("launch__registers_per_thread"),
           "stall_barrier_per_issue": vals.get("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
           "stall_long_scoreboard_per_issue": vals.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
           "stall_short_scoreboard_per_issue": vals.get("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio")}
    return out


def parity_margins():
    """worst margins observed by tests/test_gpu_long_parity.py on the B200 (committed copy under profiles/)"""
    path = os.path.join(ROOT, "profiles", "r02_parity_margins.json")
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    out = {"source": "profiles/r02_parity_margins.json (tests/test_gpu_long_parity.py, 16 perturbed members x (3+50) years "
                     "vs the oracle; config 2 vs the reference-derived fixture)"}
    for k, v in d.items():
        out[k] = {kk: vv for kk, vv in v.items() if kk != "per_member"}
    return out


WORKLOAD = "configs[2]: 1024-member perturbed-parameter/CO2 ensemble per GPU, 96x48, synthetic S0 forcing"


def workload_config(members_per_gpu: int, shared_physics: bool = False) -> dict:
    """the `config` both arms print, key for key: what is computed, not how"""
    return {"workload": WORKLOAD, "members_per_gpu": members_per_gpu, "grid": "96x48",
            "step": "one simulated year (730 steps, 12 month-end outputs x 5 fields) for every member",
            "physics": "shared (CO2-only)" if shared_physics else "perturbed per member (campaign.perturbed_member)"}


def member_physics(m: int, default_physics=None):
    """SURVEY.md 8d config 3 draws (greb_b200/campaign.py)."""
    from greb_b200 import campaign
    return campaign.perturbed_member(m)


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        if shutil.which("nvidia-smi") is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's CPU path = one `./greb` process per host core (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_baseline(years: int, forcing=None, cores: int | None = None) -> dict:
    from oracle import oracle as om
    from greb_b200 import synth
    om.build()
    cores = cores or os.cpu_count() or 1
    tmp = tempfile.mkdtemp(prefix="greb_cpu_")
    try:
        f = forcing if forcing is not None else synth.cached_forcing()
        f.write(os.path.join(tmp, "input"))
        os.makedirs(os.path.join(tmp, "output"), exist_ok=True)
        procs = []
        for c in range(cores):
            nml = os.path.join(tmp, f"nml_{c}")
            p, co2 = member_physics(c)            # core c integrates member c of the same perturbed ensemble
            phys = "".join(f"{k} = {getattr(p, k)!r}\n" for k in ("kappa", "ct_sens", "ce", "co_turb", "a_cloud", "da_ice"))
            with open(nml, "w") as fh:
                fh.write("&PHYSICS_PAR\n%s/\n&NUMERICS_PAR\ntime_flux = 0\ntime_scnr = %d\n/\n&DIAGNOSTICS_PAR\n"
                         "ens_id = \"%d\"\n/\n&CO2_PAR\nco2_ppm = %r\n/\n" % (phys, years, c, co2))
        t0 = time.perf_counter()
        for c in range(cores):
            cmd = [om.CLI, os.path.join(tmp, f"nml_{c}"), "--no-output", "--time"]
            if shutil.which("taskset"):
                cmd = ["taskset", "-c", str(c)] + cmd
            procs.append(subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, text=True))
        scen = []
        for p in procs:
            out = p.communicate()[0]
            for ln in out.splitlines():
                if ln.startswith("oracle_seconds"):
                    scen.append(float(ln.split("scenario=")[1]))
        wall = time.perf_counter() - t0
        if len(scen) != cores:
            raise RuntimeError("oracle processes failed")
        # throughput of the scenario phase with all cores busy (slowest process bounds the job)
        value = cores * years / max(scen)
        return {"value": value, "unit": "member-years/s", "cores": cores, "kind": "port",
                "sample": f"{cores} oracle processes (one per core, taskset), {years} scenario years each, members "
                          f"0..{cores - 1} of the same perturbed ensemble, S0 forcing; wall {wall:.1f}s incl. input read",
                "note": "C restatement of src/greb.f90 built -O3 -ffp-contract=off (no Fortran compiler in the image)"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    years = max(1, args.steps)
    cores = os.cpu_count() or 1
    # bounded sample: `steps` scenario years per core (~2 s per year and core)
    t0 = time.perf_counter()
    cb = cpu_baseline(years)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": "member-years/sec", "value": cb["value"], "unit": "member-years/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * cores / cb["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.members, args.shared_physics),
        "arm": {"reference_arm": "oracle port of src/greb.f90 (the reference's CPU path), one process per host core, "
                                 "each integrating one member of the ensemble for `steps` simulated years",
                "arithmetic": "IEEE fp32 in the reference's operation order, glibc libm (gcc -O3 -ffp-contract=off)"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "member-years/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=MEMBERS_PER_GPU, help="members per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--arith", default="exact", choices=["fast", "exact"],
                    help="exact (default): the reference's IEEE operation order, no FMA contraction, glibc's expf/logf "
                         "restated on the device — whole runs are bit-identical to the reference arithmetic; "
                         "fast: factored stencils + FMA + approximate division/log/exp (NOT within the 0.01 K gate "
                         "for low-CO2 perturbed members over 50 years: tests/test_gpu_long_parity.py)")
    ap.add_argument("--no-bigrid", action="store_true",
                    help="skip the strong-scaling block of BASELINE.json configs[4] (one 0.25-degree member in bands)")
    ap.add_argument("--quick", action="store_true",
                    help="kernel experiments: skip the exact-mode, single-run and CPU-baseline extras")
    ap.add_argument("--shared-physics", action="store_true",
                    help="CO2-only ensemble (config 3 (i)): one shared spin-up and correction set")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version
    # banner on file descriptor 1 whatever NCCL_DEBUG says), so descriptor 1 is pointed at stderr for the
    # whole run and the JSON line goes to a private duplicate of the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import greb_b200
    from greb_b200 import flops as fm, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    M = args.members
    K, W = args.steps, max(args.warmup, 0)
    n_years = W + K + 30      # warm-up + timed + e2e (1 + up to 12) + the other arithmetic mode (3) + slack
    forcing = synth.cached_forcing(cache_dir=os.environ.get("GREB_FORCING_CACHE", "/tmp/greb_b200_cache"))

    ens = greb_b200.Ensemble(M, device=local)
    ens.set_arithmetic(args.arith)
    ens.set_forcing(forcing)
    flops_year = 0.0
    first, last = sharding.shard_range(M * world, world, rank)   # weak scaling: M members per rank
    assert last - first == M
    for m in range(M):
        p, co2 = member_physics(first + m, greb_b200.default_physics)
        if args.shared_physics:
            p = greb_b200.default_physics()
        flops_year += fm.flops_per_member_year(p.pi, p.kappa)
        ens.set_member(m, p, np.full(n_years, co2, dtype=np.float32))
    ens.init()

    # flux-correction spin-up (one year so that the corrections the timed steps stream are real data)
    ens.spinup(1)
    spin_ms, spin_launches = ens.last_kernel_ms()
    ens.reset_scenario()

    def diag_allreduce():
        """NCCL all-reduce of ensemble sum / sum of squares of the annual global-mean Tsurf."""
        ptr, n = ens.diag_device()
        class _A:  # __cuda_array_interface__ view of the library's device buffer (no copy)
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        d = torch.as_tensor(_A(), device=f"cuda:{local}").view(-1, 2).double()
        s = torch.stack([d.sum(0), (d * d).sum(0)]).flatten()
        if world > 1:
            dist.all_reduce(s)
        return s

    def sync_all():
        ens.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up years, K timed years --------------------------
    for _ in range(W):
        ens.run_raw(1)
        diag_allreduce()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    for _ in range(K):
        ens.run_raw(1)
        ms, nl = ens.last_kernel_ms()
        ev_ms += ms
        launches += nl
        stats = diag_allreduce()
    sync_all()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ensemble mean / variance FIELDS of the last year's monthly means (untimed): reduced over the rank's members
    # on the device, then over the ranks with one NCCL reduce to rank 0 (greb_b200/sharding.py)
    f_mean, f_var, f_cnt = sharding.reduce_field_moments(ens, M, dst=0)

    # BASELINE.json configs[4] (strong scaling): one 1440x720 member in latitude bands over the same N GPUs,
    # persistent kernel with in-kernel halo exchange; 0.1 of a step timed (greb_b200/bigrid.py)
    bigrid_line = None
    if not (args.quick or args.no_bigrid or os.environ.get("GREB_BENCH_BIGRID") == "0"):
        from greb_b200 import bigrid
        bigrid_line = bigrid.bench_persistent(forcing, rank, world, local)

    t = torch.tensor([ev_ms, wall_s * 1e3], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    total_members = M * world
    value = total_members * K / (ev_ms_max / 1e3)

    # ---- end to end through the C ABI with host buffers -------------------------------------
    # Every step: H2D of every member's state from pinned memory, one simulated year, D2H of the end state and
    # of the year's diagnostic (read on the host: the step's "loss"), D2H of all 12 x 5 monthly-mean records.
    # The records of year y are copied by the library's copy stream while year y+1 runs (greb_b200_run_async +
    # greb_b200_fetch_monthly_async, two pinned record buffers) and are queued BEHIND the next step's state
    # upload: a saturated D2H stream would otherwise throttle that upload (PCIe read requests travel upstream)
    # and with it the start of the next kernel.  The timed region ends with greb_b200_wait, i.e. when the LAST
    # year's records have landed.
    e2e = None
    if not args.no_e2e:
        states = torch.empty((M, 5, 48, 96), dtype=torch.float32).pin_memory()
        monthly = [torch.empty((M, 1, 12, 5, 48, 96), dtype=torch.float32).pin_memory() for _ in range(2)]
        ens.get_states(ptr=states.data_ptr())
        ne = max(3, min(K, 12))
        ens.set_states_async(states.data_ptr())                         # the first step's input (untimed iteration 0)
        for i in range(1 + ne):
            if i == 1:
                ens.wait()
                sync_all()
                te = time.perf_counter()
            ens.run_async(1)                                            # the year's kernel
            ens.get_states_async(states.data_ptr())                     # D2H: end state (the host's copy of the result)
            ens.sync_compute()                                          # kernel + end state done
            float(diag_allreduce()[0])                                  # D2H read of the step's diagnostic
            ens.set_states_async(states.data_ptr())                     # H2D: the next step's input state ...
            ens.fetch_monthly_async(monthly[i & 1].data_ptr())          # ... then the 1.1 GB of records: the copy stream
            #                                                             starts behind the H2D and runs under the next kernel
        ens.wait()                                                      # the last year's records are on the host
        sync_all()
        dt = time.perf_counter() - te
        tt = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_members * ne / float(tt[0]), "unit": "member-years/s",
               "h2d_bytes_per_step": int(states.numel() * 4),
               "d2h_bytes_per_step": int(monthly[0].numel() * 4 + states.numel() * 4 + 32),
               "steps": ne,
               "pipeline": "records of year y copied (copy stream, pinned) while year y+1 runs; timed to the last byte"}

    if rank == 0:
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak_fp32 = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        flops_launch = flops_year  # one launch = one simulated year of this rank's members
        ms_launch = ev_ms_max / max(launches, 1)
        achieved = flops_launch / (ms_launch / 1e3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        by = fm.bytes_per_member_year(shared_corrections=args.shared_physics)
        bytes_launch = M * (by["fluxcorr_read"] + by["monthly_written"]) + by["forcing_per_gpu_year"]
        hbm_ach = bytes_launch / (ms_launch / 1e3) / 1e9
        ncu = ncu_counters(args.arith)
        issue_frac = None
        if ncu:
            # machine-side utilisation of THIS run: warp instructions per member-year (ncu capture of the same
            # build) x members / (launch time x 148 SMs x 4 schedulers x observed clock)
            issue_frac = ncu["warp_instructions_per_member_year"] * M / ((ms_launch / 1e3) * 148 * 4 * sm_mhz * 1e6)
        roofline = {
            "bound": "fp32", "achieved": achieved, "peak": peak_fp32, "unit": "TFLOP/s", "frac": achieved / peak_fp32,
            "traffic": ncu["dram_bytes_per_member_year"] * M if ncu else None,
            "traffic_note": ((f"DRAM bytes per launch from {ncu['source']} ({ncu['members_per_launch']}-member launch, "
                              f"scaled to {M} members)" if ncu else "no ncu capture committed for this build") +
                             f"; algorithmic bytes per launch = {bytes_launch:.4g}"),
            "issue_slot_frac": issue_frac,
            "kernel": "greb_member_kernel",
            "note": ("as-written reference flop count (greb_b200/flops.py, FMA=2) per launch / CUDA-event launch time; "
                     f"peak = 148 SMs x 128 lanes x 2 x {sm_mhz:.0f} MHz observed during the run; the reference "
                     "arithmetic has almost no fusable multiply-adds, so the no-FMA ceiling is peak/2 "
                     "(frac_nofma); no tensor cores (not a contraction)"),
            "frac_nofma": achieved / (peak_fp32 / 2),
            "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
        }
        line = {
            "metric": "member-years/sec", "value": value, "unit": "member-years/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ev_ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(M, args.shared_physics),
            "arm": {"l2": "inputs larger than L2 (per-member flux corrections 40 MB x members)",
                    "waves": f"{M} members on 148 SMs = {M / 148:.2f} waves of one CTA per SM "
                             f"({(1 - M / (148 * -(-M // 148))) * 100:.1f} % of the last wave idle)",
                    "arithmetic": ("fast mode (GREB_ARITH_FAST: factored stencils, FMA contraction, approximate "
                                   "division/log/exp in the column physics; config 1 and 2 and 14 of 16 perturbed "
                                   "members within 1.5e-3 K over 50 years, 2 low-CO2 members outside the 0.01 K gate)"
                                   if args.arith == "fast" else
                                   "exact mode (the reference's IEEE operation order, no FMA contraction, IEEE "
                                   "divisions, glibc's expf/logf restated on the device: 16 perturbed members x "
                                   "(3+50) years bit-identical to the oracle, tests/test_gpu_long_parity.py)")},
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "wall_ms_per_step": wall_ms_max / K,
            "spinup_member_years_per_s": M * 1 / (spin_ms / 1e3) if spin_ms > 0 else None,
            "ensemble_stats": {"sum_gmean": float(stats[0]), "sum_gmean_coslat": float(stats[1]),
                               "sumsq_gmean": float(stats[2])},
            "nonfinite_members": int(ens.flags().sum()),
            "bigrid": bigrid_line,
            "ensemble_fields": {"members": int(f_cnt), "december_tsurf_mean_K": float(f_mean[11, 0].mean()),
                                "december_tsurf_max_std_K": float(np.sqrt(f_var[11, 0].max())),
                                "note": "ensemble mean/variance of the 12 x 5 monthly fields: device reduction per rank + "
                                        "one NCCL reduce of 2 x 276,480 doubles to rank 0"},
        }
        # counters of the committed ncu --set full capture of this kernel, parsed from the file at run time
        line["roofline"]["ncu"] = ncu
        line["roofline"]["executed_fp32_note"] = (
            "frac uses the reference's AS-WRITTEN flop count (SURVEY 8d); issue_slot_frac = executed warp "
            "instructions / issue slots is the machine-side utilisation")
        line["parity"] = parity_margins()
        if world == 1 and not args.quick:
            # the same workload in the other arithmetic mode, 2 timed years
            other = "fast" if args.arith == "exact" else "exact"
            ens.set_arithmetic(other)
            ens.run_raw(1)
            ens.run_raw(2)
            ms_e, n_e = ens.last_kernel_ms()
            line[other + "_mode"] = {
                "value": M * 2 / (ms_e / 1e3), "unit": "member-years/s",
                "note": ("GREB_ARITH_FAST: factored stencils, FMA contraction, approximate division/log/exp; "
                         "secondary number — 2 of 16 perturbed members (CO2 < 300 ppm, sea-ice edge) leave the "
                         "0.01 K gate over 50 years, see parity.perturbed_fast" if other == "fast" else
                         "GREB_ARITH_EXACT: IEEE order of the reference, no FMA contraction, glibc libm restated")}
            ens.set_arithmetic(args.arith)
        if world == 1 and not args.quick:
            # BASELINE.json's second metric: single-run sim-years/s (one member, one GPU, config 1 physics)
            ens.close()
            one = greb_b200.Ensemble(1, device=local)
            one.set_arithmetic(args.arith)
            one.set_forcing(forcing)
            one.set_member(0, greb_b200.default_physics(), np.full(8, 680.0, dtype=np.float32))
            one.init()
            one.spinup(1)
            one.reset_scenario()
            one.run_raw(1)
            one.run_raw(4)
            ms1, n1 = one.last_kernel_ms()
            line["single_run_sim_years_per_s"] = 4 / (ms1 / 1e3)
            one.close()
            if not args.shared_physics:
                # BASELINE.json configs[2], variant (i): CO2-only ensemble — every member the default physics, its
                # own CO2 level; ONE physics group, i.e. one shared spin-up and one 40 MB set of flux corrections
                # that stays L2-resident instead of 41 GB streamed from HBM
                co = greb_b200.Ensemble(M, device=local)
                co.set_arithmetic(args.arith)
                co.set_forcing(forcing)
                for m in range(M):
                    co.set_member(m, greb_b200.default_physics(), np.full(6, member_physics(first + m)[1], dtype=np.float32))
                co.init()
                co.spinup(1)
                co.reset_scenario()
                co.run_raw(1)
                co.run_raw(3)
                msc, _ = co.last_kernel_ms()
                line["co2_only_ensemble"] = {"value": M * 3 / (msc / 1e3), "unit": "member-years/s", "members": M,
                                             "note": "config 3 (i): shared physics, per-member CO2; one spin-up, flux "
                                                     "corrections shared (L2-resident)"}
                co.close()
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            cb = cpu_baseline(years=6, forcing=forcing)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if getattr(ens, "h", None):
        ens.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
