"""Big-grid circulation path (include/greb_grid.h; BASELINE.json configs[4]).

One member on a grid too large for one SM (1440x720) is cut into latitude bands, one band per GPU /
process; longitude stays whole, so the periodic wrap and the polar sub-sub-steps are GPU-local.  A
sub-step needs rows k-2..k+2 (src/greb.f90:587-590, 771-780): every `s` sub-steps the ranks exchange
2*s rows with each neighbour and then advance `s` sub-steps without communication, recomputing the
shrinking halo redundantly (communication-avoiding halo).  The poles are not neighbours (no
cross-pole term, f:589-590, 756-762, 789-795), so the band chain is open-ended.

`DeviceBand` is the product (CUDA, libgreb_grid.so, fails loudly without it).  The exchange logic
(`advance`) only needs an object with `substeps / rows / halo_refreshed`, which lets the tests drive
it on the CPU over gloo with a stand-in band built on the parity oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import numpy as np

from . import lib as _lib
from . import sharding

GRID_SYMBOLS = ["greb_grid_create", "greb_grid_destroy", "greb_grid_last_error", "greb_grid_set_geometry",
                "greb_grid_set_fields", "greb_grid_substeps", "greb_grid_substeps_async", "greb_grid_sync", "greb_grid_view", "greb_grid_halo_refreshed",
                "greb_grid_get", "greb_grid_last_ms", "greb_grid_ipc_bytes", "greb_grid_ipc_export", "greb_grid_ipc_import",
                "greb_grid_run_persistent", "greb_grid_set_winds"]
_grid = None


def grid_library_path() -> str:
    return os.environ.get("GREB_GRID_LIB") or os.path.join(_lib.PKG_DIR, "libgreb_grid.so")


def load_grid_library():
    global _grid
    if _grid is not None:
        return _grid
    path = grid_library_path()
    if not os.path.exists(path):
        raise _lib.GrebError(f"{path} is missing: build it with make -C {_lib.CSRC} (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.greb_grid_create.argtypes = [C.POINTER(vp)] + [C.c_int] * 6
    L.greb_grid_destroy.argtypes = [vp]
    L.greb_grid_last_error.argtypes = [vp]
    L.greb_grid_last_error.restype = C.c_char_p
    L.greb_grid_set_geometry.argtypes = [vp, C.c_float, C.c_float, ip, fp]
    L.greb_grid_set_fields.argtypes = [vp, fp, fp, fp, fp]
    L.greb_grid_set_winds.argtypes = [vp, fp, fp]
    L.greb_grid_substeps.argtypes = [vp, C.c_int]
    L.greb_grid_substeps_async.argtypes = [vp, C.c_int]
    L.greb_grid_sync.argtypes = [vp]
    L.greb_grid_view.argtypes = [vp, C.POINTER(vp), ip, ip, ip, ip]
    L.greb_grid_halo_refreshed.argtypes = [vp]
    L.greb_grid_get.argtypes = [vp, fp]
    L.greb_grid_last_ms.argtypes = [vp, fp, ip]
    L.greb_grid_ipc_bytes.argtypes = []
    L.greb_grid_ipc_export.argtypes = [vp, vp]
    L.greb_grid_ipc_import.argtypes = [vp, C.c_int, vp]
    L.greb_grid_run_persistent.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    _grid = L
    return L


def upsample(a: np.ndarray, ny: int, nx: int, rows=None) -> np.ndarray:
    """bilinear upsample of a [48][96] cell-centred field (periodic in longitude, clamped at the poles);
    rows = (lo, hi): only those rows of the [ny][nx] result are computed (the same bits), the others are 0"""
    sy, sx = a.shape
    lo, hi = (0, ny) if rows is None else (max(int(rows[0]), 0), min(int(rows[1]), ny))
    y = (np.arange(lo, hi) + 0.5) * sy / ny - 0.5
    x = (np.arange(nx) + 0.5) * sx / nx - 0.5
    y0 = np.clip(np.floor(y).astype(int), 0, sy - 1)
    y1 = np.clip(y0 + 1, 0, sy - 1)
    fy = np.clip(y - np.floor(y), 0, 1).astype(np.float32)
    fy = np.where(np.floor(y) < 0, 0.0, fy).astype(np.float32)
    x0 = np.floor(x).astype(int) % sx
    x1 = (x0 + 1) % sx
    fx = (x - np.floor(x)).astype(np.float32)
    top = a[y0][:, x0] * (1 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (1 - fx) + a[y1][:, x1] * fx
    block = np.ascontiguousarray(top * (1 - fy[:, None]) + bot * fy[:, None], dtype=np.float32)
    if rows is None:
        return block
    out = np.zeros((ny, nx), dtype=np.float32)
    out[lo:hi] = block
    return out


def band_range(ny: int, world: int, rank: int) -> Tuple[int, int]:
    """latitude rows [k0, k1) of `rank`: contiguous bands, heights differ by at most one"""
    return sharding.shard_range(ny, world, rank)


class DeviceBand:
    """one latitude band of one member on one B200 (a handle of include/greb_grid.h)"""

    def __init__(self, nx: int, ny: int, k0: int, k1: int, s: int, device: int = 0, pi: float = 3.1416,
                 kappa: float = 8e5):
        self.L = load_grid_library()
        self.h = C.c_void_p()
        self.nx, self.ny, self.k0, self.k1, self.s, self.device = nx, ny, k0, k1, s, device
        rc = self.L.greb_grid_create(C.byref(self.h), nx, ny, k0, k1, 2 * s, device)
        if rc != 0:
            msg = self.L.greb_grid_last_error(None).decode()
            self.h = None
            raise _lib.GrebError(f"greb_grid_create failed ({rc}): {msg}")
        nsub, dt = C.c_int(), C.c_float()
        self._ck(self.L.greb_grid_set_geometry(self.h, pi, kappa, C.byref(nsub), C.byref(dt)), "greb_grid_set_geometry")
        self.nsub, self.dt_crcl = nsub.value, dt.value
        self.kernel_ms = 0.0
        self.launches = 0

    def _ck(self, rc, what):
        if rc != 0:
            raise _lib.GrebError(f"{what} failed ({rc}): {self.L.greb_grid_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.greb_grid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_fields(self, X, wz, u, v):
        """FULL global host fields [ny][nx]; the band cuts out its rows and halos"""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (X, wz, u, v)]
        assert all(a.shape == (self.ny, self.nx) for a in arrs)
        self._ck(self.L.greb_grid_set_fields(self.h, *[_lib._p(a) for a in arrs]), "greb_grid_set_fields")

    def set_winds(self, u, v):
        """the winds of another step (full global host fields); field, level and halos stay"""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (u, v)]
        assert all(a.shape == (self.ny, self.nx) for a in arrs)
        self._ck(self.L.greb_grid_set_winds(self.h, *[_lib._p(a) for a in arrs]), "greb_grid_set_winds")

    def substeps(self, n: int):
        self._ck(self.L.greb_grid_substeps(self.h, n), "greb_grid_substeps")
        ms, nl = C.c_float(), C.c_int()
        self.L.greb_grid_last_ms(self.h, C.byref(ms), C.byref(nl))
        self.kernel_ms += ms.value
        self.launches += nl.value

    def substeps_async(self, n: int):
        """queue n sub-steps on the band's stream and return; `sync` waits for them"""
        self._ck(self.L.greb_grid_substeps_async(self.h, n), "greb_grid_substeps_async")
        self._pending = True

    def sync(self):
        self._ck(self.L.greb_grid_sync(self.h), "greb_grid_sync")
        if getattr(self, "_pending", False):
            ms, nl = C.c_float(), C.c_int()
            self.L.greb_grid_last_ms(self.h, C.byref(ms), C.byref(nl))
            self.kernel_ms += ms.value
            self.launches += nl.value
            self._pending = False

    def rows(self, lo: int, hi: int):
        """torch view (no copy) of the global rows [lo, hi) of the current field buffer"""
        import torch
        ptr, kbase, nrows = C.c_void_p(), C.c_int(), C.c_int()
        self._ck(self.L.greb_grid_view(self.h, C.byref(ptr), C.byref(kbase), C.byref(nrows), None, None), "greb_grid_view")
        assert kbase.value <= lo <= hi <= kbase.value + nrows.value
        addr = ptr.value + (lo - kbase.value) * self.nx * 4

        class _A:
            __cuda_array_interface__ = {"shape": (hi - lo, self.nx), "typestr": "<f4", "data": (addr, False), "version": 2}
        return torch.as_tensor(_A(), device=f"cuda:{self.device}")

    def halo_refreshed(self):
        self._ck(self.L.greb_grid_halo_refreshed(self.h), "greb_grid_halo_refreshed")

    def ipc_export(self) -> bytes:
        """the blob a neighbour rank needs to map this band's buffers (CUDA IPC handles + geometry)"""
        buf = C.create_string_buffer(self.L.greb_grid_ipc_bytes())
        self._ck(self.L.greb_grid_ipc_export(self.h, buf), "greb_grid_ipc_export")
        return buf.raw

    def ipc_import(self, side: int, blob: bytes):
        """map the south (side 0) / north (side 1) neighbour's buffers"""
        buf = C.create_string_buffer(blob, len(blob))
        self._ck(self.L.greb_grid_ipc_import(self.h, side, buf), "greb_grid_ipc_import")

    def get(self) -> np.ndarray:
        out = np.zeros((self.k1 - self.k0, self.nx), dtype=np.float32)
        self._ck(self.L.greb_grid_get(self.h, _lib._p(out)), "greb_grid_get")
        return out


def exchange_halos(band, rank: int, world: int, group=None):
    """Refresh the 2*s halo rows on both inner sides from the neighbours' own rows
    (torch.distributed point-to-point: NCCL between GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    h = 2 * band.s
    ops, keep = [], []
    if rank > 0:                                            # south neighbour
        send = band.rows(band.k0, band.k0 + h)
        recv = band.rows(band.k0 - h, band.k0)
        ops += [dist.P2POp(dist.isend, send, rank - 1, group), dist.P2POp(dist.irecv, recv, rank - 1, group)]
        keep += [send, recv]
    if rank < world - 1:                                    # north neighbour
        send = band.rows(band.k1 - h, band.k1)
        recv = band.rows(band.k1, band.k1 + h)
        ops += [dist.P2POp(dist.isend, send, rank + 1, group), dist.P2POp(dist.irecv, recv, rank + 1, group)]
        keep += [send, recv]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if keep[0].is_cuda:
            # NCCL runs on torch's stream, the kernels on the handle's: the received rows must have
            # landed before the next sub-step is launched
            import torch
            torch.cuda.current_stream(keep[0].device).synchronize()
    band.halo_refreshed()


def advance(band, n_substeps: int, rank: int = 0, world: int = 1, group=None) -> int:
    """`n_substeps` circulation sub-steps of the whole domain, exchanging halos every band.s
    sub-steps.  Returns the number of exchanges."""
    if world > 1 and band.k1 - band.k0 < 2 * band.s:
        raise ValueError(f"band of {band.k1 - band.k0} rows is thinner than the 2*s = {2 * band.s} halo rows")
    done = exchanges = 0
    while done < n_substeps:
        n = min(band.s, n_substeps - done)
        if world > 1:
            exchange_halos(band, rank, world, group)
            exchanges += 1
        else:
            band.halo_refreshed()
        band.substeps(n)
        done += n
    return exchanges


def advance_overlapped(bands, n_substeps: int, rank: int = 0, world: int = 1, group=None) -> int:
    """The same for several independent fields of one band (air temperature and humidity of a step,
    src/greb.f90:299-304) with the exchange of one field hidden behind the sub-steps of the others:
    while field A's `s` sub-steps run on its stream, the host waits for field B's previous batch and
    exchanges B's halos.  Returns the number of exchanges."""
    s = bands[0].s
    if world > 1 and any(b.k1 - b.k0 < 2 * b.s or b.s != s for b in bands):
        raise ValueError("bands must share s and be at least 2*s rows high")
    done = exchanges = 0
    while done < n_substeps:
        n = min(s, n_substeps - done)
        for b in bands:
            b.sync()                                        # its previous batch is complete
            if world > 1:
                exchange_halos(b, rank, world, group)
                exchanges += 1
            else:
                b.halo_refreshed()
            b.substeps_async(n)
        done += n
    for b in bands:
        b.sync()
    return exchanges


class PersistentGroup:
    """The persistent path (include/greb_grid.h greb_grid_run_persistent): the fields of one band — up to two
    DeviceBands with the same rows, created with s = 1, i.e. 2 halo rows — advance together in ONE
    cooperative launch per call; halo rows travel GPU to GPU inside the kernel (peer stores + flags), the
    host only swaps the CUDA IPC handles once."""

    def __init__(self, bands, rank: int = 0, world: int = 1, group=None):
        self.bands, self.rank, self.world = list(bands), rank, world
        if not 1 <= len(self.bands) <= 2:
            raise ValueError("a group holds one or two fields")
        self.L = self.bands[0].L
        self.kernel_ms = 0.0
        self.launches = 0
        if world > 1:
            connect_neighbours(self.bands, rank, world, group)

    def advance(self, n: int):
        arr = (C.c_void_p * len(self.bands))(*[b.h for b in self.bands])
        rc = self.L.greb_grid_run_persistent(arr, len(self.bands), n)
        self.bands[0]._ck(rc, "greb_grid_run_persistent")
        ms, nl = C.c_float(), C.c_int()
        self.L.greb_grid_last_ms(self.bands[0].h, C.byref(ms), C.byref(nl))
        self.kernel_ms += ms.value
        self.launches += nl.value


def connect_neighbours(bands, rank: int, world: int, group=None):
    """every rank publishes one IPC blob per field; rank r imports those of r-1 (south) and r+1 (north)"""
    import torch.distributed as dist
    mine = [b.ipc_export() for b in bands]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    for f, b in enumerate(bands):
        if rank > 0:
            b.ipc_import(0, everyone[rank - 1][f])
        if rank < world - 1:
            b.ipc_import(1, everyone[rank + 1][f])
    dist.barrier(group=group)


def quarter_degree_fields(forcing, nx: int = 1440, ny: int = 720, ityr: int = 200):
    """the synthetic 0.25-degree workload of BASELINE.json configs[4]: step `ityr` of the S0 climatologies
    bilinearly upsampled -> {"Ta": (X, wz_air), "q": (X, wz_vapor)}, u, v"""
    topo = upsample(forcing.z_topo, ny, nx)
    fld = {"Ta": (upsample(forcing.tclim[ityr - 1], ny, nx), np.exp(-topo / np.float32(8400.0)).astype(np.float32)),
           "q": (upsample(forcing.qclim[ityr - 1], ny, nx), np.exp(-topo / np.float32(5000.0)).astype(np.float32))}
    return fld, upsample(forcing.uclim[ityr - 1], ny, nx), upsample(forcing.vclim[ityr - 1], ny, nx)


def bench_persistent(forcing, rank: int, world: int, device: int, substeps: int = 540, nx: int = 1440, ny: int = 720):
    """Strong-scaling measurement of the persistent path for bench.py: both fields of one 0.25-degree member,
    latitude bands over `world` GPUs, `substeps` sub-steps per field timed after a 48-sub-step warm-up.
    EVERY rank executes the same sequence of collectives whatever happens locally (a local failure is
    carried in `ok` and reported, never turned into a missing collective).  Returns a dict on every rank."""
    import time
    import torch
    import torch.distributed as dist
    multi = world > 1 and dist.is_available() and dist.is_initialized()
    ok, err, bands, grp = True, "", [], None

    def sync():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()

    try:
        fld, u, v = quarter_degree_fields(forcing, nx, ny)
        k0, k1 = band_range(ny, world, rank)
        for name in ("Ta", "q"):
            b = DeviceBand(nx, ny, k0, k1, 1, device=device)
            b.set_fields(fld[name][0], fld[name][1], u, v)
            bands.append(b)
        blobs = [b.ipc_export() for b in bands]
    except Exception as e:  # noqa: BLE001
        ok, err, blobs = False, f"setup: {e}", [b"", b""]
    everyone = [blobs]
    if multi:
        everyone = [None] * world
        dist.all_gather_object(everyone, blobs)
    try:
        if ok and multi:
            for f, b in enumerate(bands):
                if rank > 0:
                    b.ipc_import(0, everyone[rank - 1][f])
                if rank < world - 1:
                    b.ipc_import(1, everyone[rank + 1][f])
        if ok:
            grp = PersistentGroup(bands)
    except Exception as e:  # noqa: BLE001
        ok, err = False, f"ipc: {e}"
    flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{device}")
    if multi:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    all_ok = bool(flag.item() > 0.5)
    wall = kms = 0.0
    nsub = bands[0].nsub if bands else 0
    if all_ok:
        sync()                                              # every rank has set its fields and cleared its flags
        try:
            grp.advance(48)
            grp.kernel_ms = 0.0
        except Exception as e:  # noqa: BLE001
            ok, err = False, f"warm-up: {e}"
        sync()
        t0 = time.perf_counter()
        try:
            if ok:
                grp.advance(substeps)
        except Exception as e:  # noqa: BLE001
            ok, err = False, f"run: {e}"
        sync()
        wall = time.perf_counter() - t0
        kms = grp.kernel_ms if grp else 0.0
    t = torch.tensor([wall, kms, 1.0 if ok else 0.0], dtype=torch.float64, device=f"cuda:{device}")
    if multi:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        wall, kms, okf = float(tm[0]), float(tm[1]), float(t[2])
    else:
        okf = float(t[2])
    sync()
    for b in bands:
        b.close()
    if not all_ok or okf < 0.5:
        return {"error": err or "another rank failed"}
    steps = substeps / nsub
    return {"metric": "12-hour steps/s (both circulations of one 0.25-degree member)", "value": steps / (kms / 1e3),
            "unit": "steps/s", "scaling": "strong", "n_gpus": world, "grid": f"{nx}x{ny}", "substeps_timed_per_field": substeps,
            "substeps_per_circulation": nsub, "us_per_substep": 1e3 * kms / substeps, "wall_steps_per_s": steps / wall,
            "path": "persistent cooperative kernel; halo rows pushed GPU to GPU inside the kernel (CUDA IPC + flags)"}


# ------------------------------------------------------------------------------------------------
#   A WHOLE 12-hour step on a grid of any size: column physics on tiles + the two circulations
# ------------------------------------------------------------------------------------------------
TILE = _lib.NC                       # 4,608 cells, laid out like an ensemble member
GF_COUNT, GS_COUNT, GA_COUNT = 10, 5, 6


class BigStep:
    """One latitude band [k0, k1) of one member on a grid xdim x ydim (xdim a multiple of 96), stepped through
    whole 12-hour steps (src/greb.f90:239-308): the cell-local physics runs on tiles of 4,608 consecutive cells
    through greb_b200_tile_phase (the member kernel's own column functions), the two circulations through the
    band's DeviceBands (persistent path; the only exchange between ranks).  `static` = dict of FULL global
    host fields z_topo, glacier, mld_max (for z_ocean = 3 * max over the year, f:179-183); `step_forcing(it)`
    returns the FULL global host fields of step `it`: tclim, swet, u, v, mld, mld_prev, cld and solar [ydim].
    Flux corrections are zero unless set in `self.corr` ([ntiles][3][4608], device)."""

    def __init__(self, nx, ny, static, state0, physics=None, rank=0, world=1, device=0, arith="exact", group=None):
        import torch
        if nx % _lib.XD:
            raise ValueError("xdim must be a multiple of 96 (one solar value per 96-cell segment)")
        self.t = torch
        self.nx, self.ny, self.rank, self.world, self.device = nx, ny, rank, world, device
        self.k0, self.k1 = band_range(ny, world, rank)
        self.ncell = (self.k1 - self.k0) * nx
        self.nt = -(-self.ncell // TILE)
        self.p = physics if physics is not None else _lib.default_physics()
        self.arith = {"exact": 0, "fast": 1}[arith]
        self.L = _lib.load_library()
        dev = f"cuda:{device}"
        self.dev = dev
        sl = slice(self.k0, self.k1)
        f32 = np.float32
        z = np.ascontiguousarray(static["z_topo"][sl], dtype=f32)
        gl = np.ascontiguousarray(static["glacier"][sl], dtype=f32)
        mask = ((z >= 0) * 1 + (z < 0) * 2 + (gl > 0.5) * 4 + (z > 0) * 8).astype(np.int32)
        zoc = (f32(3.0) * np.ascontiguousarray(static["mld_max"][sl], dtype=f32)).astype(f32)
        wz = np.stack([_lib.wz_field(z, self.p.z_air), _lib.wz_field(z, self.p.z_vapor)])   # glibc expf, like greb_setup
        self.z_topo = z
        self.zoc = zoc
        self.mask = self._tiles(mask.astype(np.int32), dtype=torch.int32)
        self.z_ocean = self._tiles(zoc)
        self.wz = torch.stack([self._tiles(wz[0]), self._tiles(wz[1])], dim=1).contiguous()       # [nt][2][TILE]
        self.state = torch.stack([self._tiles(np.ascontiguousarray(state0[n][sl], dtype=f32))
                                  for n in ("Ts", "Ta", "To", "q", "cap_surf")], dim=1).contiguous()   # [nt][5][TILE]
        self.acc = torch.zeros((self.nt, GA_COUNT, TILE), dtype=torch.float32, device=dev)
        self.stash = torch.zeros((self.nt, 2, TILE), dtype=torch.float32, device=dev)
        self.corr = torch.zeros((self.nt, 3, TILE), dtype=torch.float32, device=dev)
        self.X = torch.zeros((self.nt, TILE), dtype=torch.float32, device=dev)
        self.forc = torch.zeros((self.nt, GF_COUNT, TILE), dtype=torch.float32, device=dev)
        self.solar = torch.zeros((self.nt, _lib.YD), dtype=torch.float32, device=dev)
        seg_row = (np.arange(self.nt * TILE) // _lib.XD * _lib.XD // nx).reshape(self.nt, _lib.YD, -1)[:, :, 0]
        self.seg_row = np.minimum(seg_row, self.k1 - self.k0 - 1)             # band row of every 96-cell segment
        # the two circulating fields of the band (persistent path: 2 halo rows)
        full = lambda a: np.ascontiguousarray(a, dtype=f32)
        zfull = full(static["z_topo"])
        wzf = [_lib.wz_field(zfull, self.p.z_air), _lib.wz_field(zfull, self.p.z_vapor)]
        zero = np.zeros((ny, nx), dtype=f32)
        self.bands = []
        for i, name in enumerate(("Ta", "q")):
            b = DeviceBand(nx, ny, self.k0, self.k1, 1, device=device, pi=self.p.pi, kappa=self.p.kappa)
            b.set_fields(full(state0[name]), wzf[i], zero, zero)
            self.bands.append(b)
        self.grp = PersistentGroup(self.bands, rank, world, group)
        self.nsub = self.bands[0].nsub
        self.kernel_ms = 0.0

    def close(self):
        for b in getattr(self, "bands", []):
            b.close()
        self.bands = []

    # natural (row-major) band field <-> tile layout: cell c = tile c // 4608, slot c % 4608; padded to whole tiles
    def _tiles(self, a, dtype=None):
        t = self.t
        flat = np.ascontiguousarray(a).reshape(-1)
        pad = self.nt * TILE - flat.size
        if pad:
            flat = np.concatenate([flat, np.repeat(flat[-1:], pad)])
        x = t.from_numpy(flat.reshape(self.nt, TILE).copy()).to(self.dev)
        return x if dtype is None else x.to(dtype)

    def field(self, name) -> np.ndarray:
        """the band's rows of a state field, host [k1-k0][xdim]"""
        i = {"Ts": 0, "Ta": 1, "To": 2, "q": 3, "cap_surf": 4}[name]
        return self.state[:, i, :].reshape(-1)[:self.ncell].cpu().numpy().reshape(self.k1 - self.k0, self.nx)

    def _phase(self, phase, co2):
        p = lambda x: C.c_void_p(x.data_ptr())
        rc = self.L.greb_b200_tile_phase(self.device, self.arith, phase, self.nt, C.byref(self.p), C.c_float(co2),
                                         p(self.forc), p(self.solar), p(self.mask), p(self.z_ocean), p(self.wz),
                                         p(self.corr), p(self.state), p(self.acc), p(self.stash), p(self.X))
        if rc != 0:
            raise _lib.GrebError(f"greb_b200_tile_phase failed ({rc}): {self.L.greb_b200_last_error(None).decode()}")

    def _circulate(self, band_index, state_index):
        """X = the state field -> band buffer (own rows + the neighbours' 2 halo rows), nsub sub-steps, back"""
        t = self.t
        b = self.bands[band_index]
        rows = self.k1 - self.k0
        nat = self.state[:, state_index, :].reshape(-1)[:self.ncell].reshape(rows, self.nx)
        b.rows(self.k0, self.k1).copy_(nat)
        t.cuda.synchronize()
        if self.world > 1:
            exchange_halos(b, self.rank, self.world)          # the field's halo rows at the start of the circulation
            import torch.distributed as dist
            dist.barrier()
        return b

    def stage_forcing(self, f):
        """One step's forcing -> a staged set on the device: tiles (host arithmetic = greb_setup.cpp's, fp32 IEEE),
        solar per segment, and the winds' host fields for the bands.  Touches nothing the running step uses (own
        CUDA stream, pinned staging buffer, two device sets used in turn), so a worker thread may stage step
        it+1 while step it runs (`run`)."""
        t, f32 = self.t, np.float32
        sl = slice(self.k0, self.k1)
        u, v = (np.ascontiguousarray(f[n][sl], dtype=f32) for n in ("u", "v"))
        mld, mldp = (np.ascontiguousarray(f[n][sl], dtype=f32) for n in ("mld", "mld_prev"))
        aw = np.sqrt(u * u + v * v).astype(f32)                                        # f:452
        aw = np.where(self.z_topo > 0, np.sqrt(aw * aw + f32(2.0) * f32(2.0)).astype(f32), aw)   # f:453
        aw = np.where(self.z_topo < 0, np.sqrt(aw * aw + f32(3.0) * f32(3.0)).astype(f32), aw)   # f:454
        dmld = (mld - mldp).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            rdeep = (dmld / (self.zoc - mld)).astype(f32)
            rmix = (dmld / mld).astype(f32)
        dtrad = (f32(-0.16) * np.ascontiguousarray(f["tclim"][sl], dtype=f32) - f32(5.0)).astype(f32)   # f:176
        fields = [u, v, np.ascontiguousarray(f["cld"][sl], dtype=f32), dtrad, np.ascontiguousarray(f["swet"][sl], dtype=f32),
                  aw, mld, dmld, rdeep, rmix]
        if not hasattr(self, "_stage"):
            with t.cuda.device(self.device):
                self._stage_stream = t.cuda.Stream()
            self._stage = [{"forc": t.empty_like(self.forc), "solar": t.empty_like(self.solar),
                            "h_forc": t.empty(self.forc.shape, dtype=t.float32).pin_memory(),
                            "h_solar": t.empty(self.solar.shape, dtype=t.float32).pin_memory(),
                            "event": t.cuda.Event()} for _ in range(2)]
            self._stage_seq = 0
        st = self._stage[self._stage_seq % 2]
        self._stage_seq += 1
        st["event"].synchronize()                 # the previous upload into this set (its step is long over)
        hf = st["h_forc"].numpy()
        for j, a in enumerate(fields):            # natural band order -> tiles, the tail padded with the last cell
            dst = hf[:, j, :]
            src = a.reshape(-1)
            full = src.size // TILE
            dst[:full] = src[:full * TILE].reshape(full, TILE)
            if full < self.nt:
                rest = src.size - full * TILE
                dst[full, :rest] = src[full * TILE:]
                dst[full, rest:] = src[-1]
        sol = np.ascontiguousarray(f["solar"], dtype=f32)[self.k0:self.k1]
        st["h_solar"].numpy()[...] = sol[self.seg_row]
        with t.cuda.device(self.device), t.cuda.stream(self._stage_stream):
            st["forc"].copy_(st["h_forc"], non_blocking=True)
            st["solar"].copy_(st["h_solar"], non_blocking=True)
            st["event"].record()
        return {"_staged": st, "u": f["u"], "v": f["v"]}

    def _load_forcing(self, f):
        """make a step's forcing (host fields or a staged set) the current one"""
        sf = f if "_staged" in f else self.stage_forcing(f)
        st = sf["_staged"]
        st["event"].synchronize()
        self.forc, self.solar = st["forc"], st["solar"]
        for b in self.bands:
            b.set_winds(sf["u"], sf["v"])

    def step(self, it, forcing_of_step, co2):
        """one time_loop call (f:239-274) with step counter `it` (1-based; the caller's calendar picks the forcing);
        `forcing_of_step` = the FULL global host fields of the step, or what stage_forcing returned for them"""
        self._load_forcing(forcing_of_step)
        self._phase(0, co2)
        # circulation(Ta) and circulation(q) (f:301, f:303) both start from the fields of the step's beginning and
        # do not depend on each other: they run as ONE two-field persistent launch
        for bi, si in ((0, 1), (1, 3)):
            self._circulate(bi, si)
        self.grp.advance(self.nsub)
        self.kernel_ms += self.grp.kernel_ms
        self.grp.kernel_ms = 0.0
        for bi, ph in ((0, 1), (1, 2)):
            self.X.reshape(-1)[:self.ncell].copy_(self.bands[bi].rows(self.k0, self.k1).reshape(-1))
            self._phase(ph, co2)

    def run(self, it0, nsteps, step_forcing, co2=680.0, prefetch=True):
        """steps it0 .. it0+nsteps-1; with `prefetch` a worker thread prepares and uploads the forcing of step
        it+1 (step_forcing(it+1) and stage_forcing) while step it runs on the device — the host arithmetic
        and the H2D copy leave the critical path, the results are the same bits"""
        co2_of = co2 if callable(co2) else (lambda it: co2)
        if not prefetch or nsteps < 2:
            for it in range(it0, it0 + nsteps):
                self.step(it, step_forcing(it), co2_of(it))
            return
        from concurrent.futures import ThreadPoolExecutor
        job = lambda it: self.stage_forcing(step_forcing(it))
        with ThreadPoolExecutor(1) as pool:
            nxt = pool.submit(job, it0)
            for it in range(it0, it0 + nsteps):
                staged = nxt.result()
                if it + 1 < it0 + nsteps:
                    nxt = pool.submit(job, it + 1)
                self.step(it, staged, co2_of(it))


def s0_static_and_forcing(forcing, nx, ny, rows=None):
    """BigStep inputs from the synthetic S0 set (bilinearly upsampled when the grid is not 96x48): the static
    dict and a function it -> the fields of step it.  rows = (lo, hi): the step fields are filled in only for
    those rows (a rank's band plus the 2 halo rows the winds need) — 1/N of the host work on N ranks."""
    same = (nx, ny) == (_lib.XD, _lib.YD)
    up = (lambda a: np.ascontiguousarray(a, dtype=np.float32)) if same else (lambda a: upsample(a, ny, nx))
    ups = up if (same or rows is None) else (lambda a: upsample(a, ny, nx, rows))
    static = {"z_topo": up(forcing.z_topo), "glacier": up(forcing.glacier), "mld_max": up(forcing.mldclim.max(axis=0))}
    lat_src = (np.arange(_lib.YD) + 0.5) / _lib.YD
    lat_dst = (np.arange(ny) + 0.5) / ny

    def step_forcing(it):
        n = (it - 1) % _lib.NT
        npv = n - 1 if n > 0 else _lib.NT - 1
        sol = forcing.sw_solar[n] if ny == _lib.YD else np.interp(lat_dst, lat_src, forcing.sw_solar[n]).astype(np.float32)
        return {"tclim": ups(forcing.tclim[n]), "swet": ups(forcing.swetclim[n]), "u": ups(forcing.uclim[n]),
                "v": ups(forcing.vclim[n]), "mld": ups(forcing.mldclim[n]), "mld_prev": ups(forcing.mldclim[npv]),
                "cld": ups(forcing.cldclim[n]), "solar": np.ascontiguousarray(sol, dtype=np.float32)}
    return static, step_forcing
