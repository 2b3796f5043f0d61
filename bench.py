"""bench.py — float64 particle-steps/s of the skeletor particle hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

One STEP = one pass of the hot path over all particles of the workload:
    ions.push(E, B, dt)  ->  sources.deposit(ions)  ->  sources.add_guards()
    ->  sources.copy_guards()
(the loop body of reference tests/test_ionacoustic.py:160-178 without Ohm;
SURVEY.md §8d), i.e. gather + Boris push + fused boundary epilogue + migration +
tile sort + deposit + guard cells.  One particle-step = one particle through that.

Workload (BASELINE.json config 5, the one the metric/target is quoted on): uniform
Maxwellian plasma, 2048 x 2048 grid x 256 particles/cell = 1.07e9 particles, CIC,
vt*dt/dx = 0.1, smooth E ~ 0.01, B = z-hat, float64, synthetic (torch.Generator
seed 1234 + rank).  STRONG scaling: the grid is split into N y-slabs, one per GPU.

JSON keys beyond the base contract:
  roofline     dominant kernel vs measured HBM peak (MEASURED_PEAKS.json), from CUDA
               events around that kernel in an instrumented pass (not the timed one)
  kernels      the same for every kernel of the step
  cpu_baseline the unmodified reference's compiled kernels (oracle/_ref) on the
               box's host cores, bounded sample (rank 0, N = 1 only)
  e2e          same step driven with HOST field buffers: E,B copied H2D from pinned
               memory and the sources read back D2H every step
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "float64 particle-steps/sec"
UNIT = "particle-steps/s"
NX = NY = 2048
PPC = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=NX)
    ap.add_argument("--ny", type=int, default=NY)
    ap.add_argument("--ppc", type=int, default=PPC)
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--layout", default="gapped", choices=["gapped", "dense"],
                    help="particle layout: per-cell slot ranges with slack (default) or "
                         "dense arrays re-ordered by the tile sort every step")
    ap.add_argument("--perturbed", action="store_true",
                    help="5 %% sinusoidal density contrast along x (SURVEY.md 8d)")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: ny grows with the GPU count (ny rows PER GPU)")
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json config: 5 (default, the one the metric is quoted on) "
                         "uniform plasma push+deposit; 2 Landau loop incl. Ohm, 1024^2 x 64 "
                         "ppc; 3 Horowitz iterate, 2048^2 x 128 ppc, Hall Ohm + Faraday; 4 "
                         "shearing sheet (push_modified + sheared guards), 2048^2 x 256 ppc")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the oracle parity checks after the timed region")
    return ap.parse_args()


def workload_config(a, extra=None):
    cfg = {"workload": "uniform Maxwellian plasma %dx%d grid x %d ppc (BASELINE config 5), "
                       "push+deposit+add_guards+copy_guards per step" % (a.nx, a.ny, a.ppc),
           "grid": [a.nx, a.ny], "ppc": a.ppc, "particles": a.nx*a.ny*a.ppc,
           "interpolation": "CIC" if a.order == 1 else "TSC",
           "plasma": "perturbed (5 % density contrast)" if a.perturbed else "uniform",
           "decomposition": "y-slabs, 1 per GPU", "vt_dt_over_dx": 0.1,
           "l2_policy": "inputs larger than L2 (>= 5 GB of particle data per GPU)"}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------
def reference_arm(a):
    """--impl reference: the reference's own CPU implementation on the host cores,
    every step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_driver
    rows = 8
    r = ref_driver.measure(nx=a.nx, rows=rows, ppc=a.ppc, steps=a.steps,
                           warmup=min(a.warmup, 1))
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(a, {"sample": r["sample"]}),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                             "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: an NVML
    polling thread (5 ms period; nvidia-smi -lms needs ~100 ms to deliver its first
    line, longer than the timed region of a multi-GPU run), nvidia-smi as fall-back."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.sm, self.mx, self.pw, self.reasons = [], [], [], set()

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40,
                "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                self.pw.append(nv.nvmlDeviceGetPowerUsage(h)/1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, bit in bits.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            if self.stop_flag.wait(0.005):
                return

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML wants the physical one
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                out, _ = self.proc.communicate()
            for ln in out.splitlines():
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 8:
                    continue
                try:
                    self.sm.append(float(f[1])); self.mx.append(float(f[2]))
                    self.pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(self.NAMES, f[4:8]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
        sm, mx, pw = self.sm, self.mx, self.pw
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(self.reasons),
                "source": "nvml" if self.thread is not None else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------
def parity_checks(sk, ions, E, B, src, dt, comm, a, n_sample=100000, band=8, modified=False):
    """Parity at BENCH SCALE and at N ranks, through the public API, against the oracle
    (checker only; runs after the timed region).

    particles: a seeded sample of ~n_sample live particle rows per rank is pushed on the
      host with oracle.push + periodic_y wrap + periodic_x (reference
      particles.py:159-188: push, periodic_y, periodic_x); after ONE more ions.push on the
      GPUs every expected row must be present, bit for bit and exactly once, in the union
      of the ranks' live particles (the reference's own "N ranks == 1 rank" invariant,
      tests/test_skeletor.py:142-150, at the benchmark's size).
    sources: all particles whose stencil touches a band of `band` array rows (the top
      rows of the slab incl. one guard row) are deposited with oracle.deposit
      (deposit.pyx:6-34) and normalised (sources.py:52-63); compared with what
      sources.deposit(ions) left in those rows (before the guard-cell fold), relative to
      the band's maximum per component (<= 1e-12: summation order only) and, for rho,
      also cell by cell.
    """
    import numpy as np
    import torch
    from oracle import oracle as orc
    m = ions.manifold
    rank, size = comm.rank, comm.size
    dev = ions.device
    order = ions.order
    S, Omega = getattr(m, "S", None), getattr(m, "Omega", None)
    og = orc.Grid(a.nx, a.ny, rank=rank, size=size, lbx=m.lbx, lby=m.lby, Lx=m.Lx, Ly=m.Ly,
                  x0=m.x0, y0=m.y0, S=S, Omega=Omega)
    Eh = np.ascontiguousarray(np.asarray(E)).view(orc.Float3).reshape(m.myp, m.mx)
    Bh = np.ascontiguousarray(np.asarray(B)).view(orc.Float3).reshape(m.myp, m.mx)

    # ---- expected particle rows ------------------------------------------------------
    ions._dense()
    N = int(ions.N)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    idx = torch.unique(torch.randint(0, N, (min(n_sample, N),), generator=gen, device=dev))
    rows = ions._data[:, idx].t().contiguous().cpu().numpy()
    part = np.ascontiguousarray(rows).view(orc.Particle).reshape(-1).copy()
    qtmh = ions.charge/ions.mass*dt/2
    if modified:
        orc.push(part, Eh, Bh, og, order, qtmh, dt, True, float(Omega), float(S))
    else:
        orc.push(part, Eh, Bh, og, order, qtmh, dt)
    if S is not None:
        # shear_periodic_y with the particle time AFTER the push (particles.py:154-155,175)
        orc.shear_periodic_y(part, og, float(S), float(ions.time) + dt)
    # periodic_y: cppmove2 wraps on the edge ranks (pplib2.c:676-677, 692-693); with
    # |vy dt/dy| << nyp nobody moves more than one slab, so only they can cross 0 / ny
    y = part["y"]
    lo, hi = y < 0.0, y >= float(a.ny)
    y[lo] = y[lo] + float(a.ny)
    y[hi] = y[hi] - float(a.ny)
    orc.periodic_x(part, og)
    exp_rows = np.ascontiguousarray(part).view(np.float64).reshape(-1, 5)
    if size > 1:
        exp_rows = np.concatenate(comm.allgather(exp_rows))

    # ---- one more step on the GPUs -----------------------------------------------------
    (ions.push_modified if modified else ions.push)(E, B, dt)
    src.deposit(ions)
    src_raw = src.t.clone()             # normalised, guards not folded yet
    src.time = ions.time
    src.add_guards()
    src.copy_guards()

    # ---- particles: every expected row present exactly once -------------------------
    ions._dense()
    N = int(ions.N)
    M = exp_rows.shape[0]
    exp = torch.as_tensor(exp_rows, device=dev)
    ebits = exp.view(torch.int64)
    sx, perm = torch.sort(ebits[:, 0].contiguous())
    found = torch.zeros(M, dtype=torch.int64, device=dev)
    lx = ions._data[0, :N].view(torch.int64)
    CH = 1 << 26
    for s0 in range(0, N, CH):
        chunk = lx[s0:s0 + CH]
        pos = torch.searchsorted(sx, chunk).clamp_(max=M - 1)
        hit = (sx[pos] == chunk).nonzero().squeeze(1)
        if hit.numel():
            who = perm[pos[hit]]
            cand = ions._data[:, s0 + hit].t().contiguous().view(torch.int64)
            full = (cand == ebits[who]).all(dim=1)
            found.index_add_(0, who[full], torch.ones_like(who[full]))
        del pos, hit
    del lx
    if size > 1:
        found = torch.as_tensor(comm.allreduce(found.cpu().numpy(), op=sk.comm.SUM))
    n_once = int((found == 1).sum().item())
    out = {"particles_sampled": int(M), "particles_found_once": n_once,
           "particles_bitexact": n_once == M}

    # ---- sources: band of rows vs oracle.deposit -----------------------------------
    offy = (m.lby - 0.5) - m.noff + (0.5 if order == 2 else 0.0)
    r1 = m.uby + 1                       # band = array rows [r0, r1): top rows + 1 guard row
    r0 = r1 - band
    lo_i, hi_i = r0 - 1, (r1 if order == 1 else r1 + 1)     # stencil-base rows that touch it
    sel = []
    for s0 in range(0, N, CH):
        iy = (ions._data[1, s0:min(s0 + CH, N)] + offy).to(torch.int32)
        k = ((iy >= lo_i) & (iy < hi_i)).nonzero().squeeze(1)
        if k.numel():
            sel.append(ions._data[:, s0 + k].t().contiguous().cpu())
        del iy, k
    bp = torch.cat(sel).numpy() if sel else np.zeros((0, 5))
    bpart = np.ascontiguousarray(bp).view(orc.Particle).reshape(-1).copy()
    cur = og.field(orc.Float4)
    orc.deposit(bpart, cur, og, order, float(getattr(m, "S", 0.0)))
    fac = ions.charge*ions.n0*a.nx*a.ny/ions.N_global()
    ref = np.ascontiguousarray(cur).view(np.float64).reshape(m.myp, m.mx, 4)[r0:r1]*fac
    got = src_raw[r0:r1].cpu().numpy()
    scale = np.abs(ref).reshape(-1, 4).max(axis=0)
    rel = float((np.abs(got - ref).reshape(-1, 4).max(axis=0)/scale).max())
    nz = ref[..., 0] != 0.0
    rho_cell = float((np.abs(got[..., 0] - ref[..., 0])[nz]/np.abs(ref[..., 0][nz])).max())
    if size > 1:
        rel = comm.allreduce(rel, op=sk.comm.MAX)
        rho_cell = comm.allreduce(rho_cell, op=sk.comm.MAX)
    out.update({"sources_rel": rel, "sources_rho_cellwise_rel": rho_cell,
                "sources_band": "array rows [uby+1-%d, uby+1) x all columns, %d particles "
                                "per rank through oracle.deposit" % (band, bpart.shape[0])})
    return out


def b200_arm(a):
    import numpy as np
    import torch
    import skeletor_b200 as sk
    from skeletor_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm = sk.COMM_WORLD if world > 1 else sk.COMM_SELF
    rank, size = comm.rank, comm.size
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    m = sk.Manifold(a.nx, a.ny, comm, lbx=1 if a.order == 1 else 2,
                    lby=1 if a.order == 1 else 2, Lx=1.0, Ly=a.ny/a.nx)
    n_local = a.nx*m.nyp*a.ppc
    n_total = a.nx*a.ny*a.ppc
    gapped = a.layout == "gapped"
    nmax = int((1.36 if gapped else 1.05)*n_local) + 4096
    if gapped:          # + 4 slots for every empty cell of the tiles covering the grid
        nmax += 4*256*((m.mx + 15)//16)*((m.myp + 15)//16)
    nmax += nmax & 1
    ions = sk.Particles(m, nmax, charge=1.0, mass=1.0, order=a.order,
                        nbmax=max(n_local//100, 1 << 16))
    ions.gapped = gapped
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    d = ions._data
    d[0, :n_local] = torch.rand(n_local, generator=gen, device=dev, dtype=torch.float64)*a.nx
    d[1, :n_local] = m.noff + torch.rand(n_local, generator=gen, device=dev,
                                         dtype=torch.float64)*m.nyp
    d[2:5, :n_local] = torch.randn((3, n_local), generator=gen, device=dev,
                                   dtype=torch.float64)
    if a.perturbed:
        # displace x by (0.05 nx / 2 pi) sin(2 pi x / nx): ~5 % density contrast
        xx = d[0, :n_local]
        xx.add_(0.05*a.nx/(2*np.pi)*torch.sin(2*np.pi*xx/a.nx))
        xx.remainder_(float(a.nx))
    ions.N = n_local
    del d               # (no reference to the tensor: Particles swaps / releases its buffers)
    dt = 0.1*m.dx       # vt = 1  ->  vt*dt/dx = 0.1
    E = sk.Field(m, dtype=sk.Float3)
    B = sk.Field(m, dtype=sk.Float3)
    xg, yg = np.meshgrid(m.x, m.y)
    E['x'].active = 0.01*np.sin(2*np.pi*xg/m.Lx)
    E['y'].active = 0.01*np.cos(2*np.pi*yg/m.Ly)
    B['z'].active = 1.0
    E.copy_guards()
    B.copy_guards()
    src = sk.Sources(m)

    def step():
        ions.push(E, B, dt)
        src.deposit(ions)
        src.add_guards()
        src.copy_guards()

    def barrier():
        if size > 1:
            comm.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if size > 1:
            ms = comm.allreduce(ms, op=sk.comm.MAX)
        return ms

    for _ in range(max(a.warmup, 3)):
        step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    k0 = _lib.kernel_launches
    # CUDA events around the hot C-ABI calls of every timed step (no extra syncs): the
    # roofline below is the AVERAGE launch duration over the timed region
    hot = ("skb_push_gapped", "skb_push_deposit_gapped", "skb_boris_push",
           "skb_tile_sort_precounted", "skb_deposit")
    _lib.trace = {k: [] for k in hot}
    ms = timed(step, a.steps)
    live = {k: [e0.elapsed_time(e1) for e0, e1 in v] for k, v in _lib.trace.items() if v}
    _lib.trace = None
    launches = _lib.kernel_launches - k0
    clk = clocks.stop() if rank == 0 else None
    value = n_total*a.steps/(ms*1e-3)

    # ---- per-kernel roofline: instrumented pass, CUDA events around each C-ABI call
    peak, peak_src = measured_peak()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # (g) gapped layout, the default step: push on per-cell slot ranges (movers parked
    #     and re-inserted by the kernel itself) | migration + insertion of the rest |
    #     deposit | guards
    accg, accu, nlocal = None, None, 0
    fused = bool(gapped and ions._rep == "gapped" and
                 (ions.fuse_deposit is True or
                  (ions.fuse_deposit == "auto" and ions._fuse_next)))
    pname = "push_deposit" if fused else "push"
    if gapped and ions._rep == "gapped":
        accg = {pname: [], "migrate_insert": [], "sources" if fused else "deposit": [],
                "guards": []}
        for _ in range(3):
            if ions._rep != "gapped" and not ions._to_gapped():
                break
            t = [ev() for _ in range(5)]
            torch.cuda.synchronize()
            t[0].record(); cnt = ions._gap_kernel(E, B, dt, False, fused)
            t[1].record(); ions._gap_finish(cnt, fused=fused)
            t[2].record()
            src.deposit(ions)      # fused: copy of the push's grid + normalize
            t[3].record()
            src.add_guards(); src.copy_guards()
            t[4].record()
            torch.cuda.synchronize()
            nlocal = ions._gap_stats[6]
            for name, i in zip(accg, range(4)):
                accg[name].append(t[i].elapsed_time(t[i + 1]))
        # the two kernels on their own (what runs when the deposit does not follow a push)
        accu = {"push_unfused": [], "deposit_unfused": []}
        for _ in range(3):
            if ions._rep != "gapped" and not ions._to_gapped():
                break
            t = [ev() for _ in range(4)]
            torch.cuda.synchronize()
            t[0].record(); cnt = ions._gap_kernel(E, B, dt, False, False)
            t[1].record(); ions._gap_finish(cnt)
            src.t.zero_()
            t[2].record()
            _lib.call("skb_deposit", ions._c, ions.N, src.ptr, m.c, ions.order, 0.0,
                      ions._tiling_c(), torch.cuda.current_stream().cuda_stream)
            t[3].record()
            torch.cuda.synchronize()
            accu["push_unfused"].append(t[0].elapsed_time(t[1]))
            accu["deposit_unfused"].append(t[2].elapsed_time(t[3]))
        ions._dense()
    # (a) the step as it runs by default: push (+ fused boundary epilogue and sort
    #     histogram) | migration | precounted tile sort | deposit | guards
    st_ = lambda: torch.cuda.current_stream().cuda_stream
    acc = {"push": [], "migrate": [], "tile_sort": [], "deposit": [], "guards": []}
    for _ in range(3):
        t = [ev() for _ in range(6)]
        torch.cuda.synchronize()
        t[0].record(); ions._push_kernel(E, B, dt, False, count=True)
        t[1].record()
        nkeep = ions.move()
        _lib.call("skb_sort_count_rows", ions._keep.data_ptr(), nkeep, m.c, ions.order,
                  4, 4, ions._cell_counts.data_ptr(), st_())
        t[2].record(); ions.sort(precounted=True)
        t[3].record()
        src.t.zero_()
        _lib.call("skb_deposit", ions._c, ions.N, src.ptr, m.c, ions.order, 0.0,
                  ions._tiling_c(), st_())
        t[4].record()
        src.boundaries_set = False
        src.normalize(ions); src.add_guards(); src.copy_guards()
        t[5].record()
        torch.cuda.synchronize()
        for name, i in zip(acc, range(5)):
            acc[name].append(t[i].elapsed_time(t[i + 1]))
    # (b) alternatives kept in the library, for the record: push without the fused
    #     histogram + full tile sort (key pass + move); the recompute-twice fused
    #     push+sort passes
    acc2 = {"push_plain": [], "tile_sort_full": [], "push_count": [], "push_scatter": []}
    for _ in range(2):
        t = [ev() for _ in range(4)]
        torch.cuda.synchronize()
        t[0].record(); ions._push_kernel(E, B, dt, False)
        t[1].record(); ions.move()
        t[2].record(); ions.sort()
        t[3].record()
        torch.cuda.synchronize()
        acc2["push_plain"].append(t[0].elapsed_time(t[1]))
        acc2["tile_sort_full"].append(t[2].elapsed_time(t[3]))
        t = [ev() for _ in range(4)]
        n_before = ions.N
        ions.time += dt
        args, flags = ions._push_args(dt, False)
        til = ions._tiling_c()
        epi = ions._epilogue(flags, 0.0)
        cells = ions._cell_counts.data_ptr()
        t[0].record()
        _lib.call("skb_push_count", ions._c, ions.N, E.ptr, B.ptr, *args, til, epi,
                  4, 4, cells, ions.sbufl.data_ptr(), ions.sbufr.data_ptr(), ions.nbmax,
                  ions._counts.data_ptr(), comm.rank, comm.size, st_())
        t[1].record()
        nl, nr, ovf = ions._counts[:3].tolist()
        nkeep = ions._exchange(nl, nr)
        _lib.call("skb_sort_count_rows", ions._keep.data_ptr(), nkeep, args[0], ions.order,
                  4, 4, cells, st_())
        _lib.call("skb_sort_scan", cells, args[0], 4, 4, 2048, ions._block_sums.data_ptr(),
                  ions._tile_offsets.data_ptr(), ions._chunk_first.data_ptr(), st_())
        out = ions._soa(ions._alt)
        t[2].record()
        _lib.call("skb_push_scatter", ions._c, out, ions.N, E.ptr, B.ptr, *args, til, epi,
                  4, 4, cells, st_())
        t[3].record()
        _lib.call("skb_sort_scatter_rows", ions._keep.data_ptr(), nkeep, out, args[0],
                  ions.order, 4, 4, cells, st_())
        ions._data, ions._alt = ions._alt, ions._data
        ions.N = n_before - nl - nr + nkeep
        ions._n_sorted, ions._sorted = ions.N, True
        torch.cuda.synchronize()
        acc2["push_count"].append(t[0].elapsed_time(t[1]))
        acc2["push_scatter"].append(t[2].elapsed_time(t[3]))
    npart = ions.N
    cells = m.mx*m.myp
    # algorithmic bytes per launch: push 80 B/particle + E,B tiles
    # 48 B/cell; deposit 40 B/particle + 32 B/cell; precounted sort: 80 B (move), full
    # sort: + 16 B key pass; recompute passes: 40 B read / 40 B read + 40 B written
    # fused push + deposit: SURVEY.md §8d's figure for a sweep that reads and writes every
    # particle once (as push_and_deposit(update=True)): 80 B + E,B read and sources written
    alg = {"push": 80.0*npart + 48.0*cells, "deposit": 40.0*npart + 32.0*cells,
           "push_deposit": 80.0*npart + 80.0*cells,
           "push_unfused": 80.0*npart + 48.0*cells,
           "deposit_unfused": 40.0*npart + 32.0*cells,
           "tile_sort": 80.0*npart, "push_plain": 80.0*npart + 48.0*cells,
           "tile_sort_full": 96.0*npart, "push_count": 40.0*npart + 48.0*cells,
           "push_scatter": 80.0*npart + 48.0*cells}

    def summarize(accd):
        out = {}
        for name, ts in accd.items():
            tmin = min(ts)
            out[name] = {"ms": round(tmin, 4)}
            if name in alg:
                gbs = alg[name]/(tmin*1e-3)/1e9
                out[name].update({"alg_bytes": alg[name], "achieved_gbs": round(gbs, 1),
                                  "frac": round(gbs/peak, 4)})
        return out
    standalone = summarize(acc2)
    if accg and all(accg.values()):
        standalone.update({"dense_" + k: v for k, v in summarize(acc).items()})
        if accu and all(accu.values()):
            standalone.update(summarize(accu))
        # the push is rated at 80 B/particle (+ E,B tiles) although this
        # kernel also maintains the ordering (block-local movers are written, read back
        # and written again: + 80 B each, reported separately)
        kern = summarize(accg)
        kern[pname]["bytes_incl_reinsertion"] = alg[pname] + 80.0*nlocal
        candidates = (pname,) if fused else ("push", "deposit")
        tkey = {"push": "push_gapped (cell_stream_kernel, PD=0)",
                "push_deposit": "push_deposit_gapped (cell_stream_kernel, PD=3: push + "
                                "full-step deposit in one sweep)"}
    else:
        kern = summarize(acc)
        candidates, tkey = ("push", "tile_sort", "deposit"), {}
    step_ms = sum(v["ms"] for v in kern.values())
    for v in kern.values():
        v["share"] = round(v["ms"]/step_ms, 4)
    dom = max(candidates, key=lambda k: kern[k]["ms"])
    # live numbers of the timed region: average duration of the C-ABI call that
    # launches the kernel (skb_deposit is also called for the handful of leftover
    # particles of the gapped layout: keep the big launches only)
    entry = {"push": "skb_push_gapped" if accg else "skb_boris_push",
             "push_deposit": "skb_push_deposit_gapped",
             "tile_sort": "skb_tile_sort_precounted", "deposit": "skb_deposit"}
    for name, ep in entry.items():
        ts = [t for t in live.get(ep, []) if t > 0.2*max(live[ep])]
        if ts and name in kern:
            avg = sum(ts)/len(ts)
            kern[name].update({"live_ms": round(avg, 4), "live_launches": len(ts),
                               "live_frac": round(alg[name]/(avg*1e-3)/1e9/peak, 4)})
    if "live_ms" in kern[dom]:
        dom_ms, how = kern[dom]["live_ms"], ("CUDA events around the launch in every "
                                             "timed step, average")
    else:
        dom_ms, how = kern[dom]["ms"], "instrumented pass after the timed region, best of 3"
    dom_gbs = alg[dom]/(dom_ms*1e-3)/1e9
    # bytes one particle-step has to move: 120 (push 80 + deposit 40)
    # when the two are separate kernels, 80 when the deposit is fused into the push
    step_bytes = 80.0 if fused else 120.0
    roofline = {"kernel": tkey.get(dom, dom), "bound": "hbm",
                "achieved": round(dom_gbs, 1),
                "peak": peak, "unit": "GB/s", "frac": round(dom_gbs/peak, 4),
                "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_launch": alg[dom],
                "alg_bytes_per_particle": 80.0 if dom in ("push", "push_deposit") else
                (40.0 if dom == "deposit" else None),
                "ms_per_launch": dom_ms, "timing": how,
                "frac_best_isolated": kern[dom]["frac"],
                # whole step against the bytes it has to move per particle-step
                # (ordering, migration and halo traffic are overhead)
                "step_bytes_per_particle": step_bytes,
                "step_frac": round(step_bytes*npart/(ms/a.steps*1e-3)/1e9/peak, 4),
                "step_frac_unfused_accounting":
                    round(120.0*npart/(ms/a.steps*1e-3)/1e9/peak, 4)}
    # the other target kernels of the north star, each on its own bytes
    others = {}
    for name in ("push_unfused", "deposit_unfused", "deposit", "push"):
        src_d = kern if name in kern else standalone
        if name in src_d and name != dom and "frac" in src_d[name]:
            others[name] = {k: src_d[name][k] for k in
                            ("ms", "achieved_gbs", "frac", "live_ms", "live_frac")
                            if k in src_d[name]}
    roofline["others"] = others
    # DRAM traffic: bytes per particle of one `ncu --set full` capture of this kernel at
    # full size (profiles/traffic_r02.json names the capture), scaled by the particles
    # this launch processed
    tr = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tr) and a.order == 1:          # (the captures are of the CIC kernels)
        try:
            rec = json.load(open(tr)).get(dom)
            if rec:
                roofline["traffic"] = rec["dram_bytes_per_particle"]*npart
                roofline["traffic_source"] = (
                    "ncu dram__bytes_read.sum + dram__bytes_write.sum of %s, %.1f B per "
                    "particle, scaled by the %d particles of this launch" % (
                        rec["capture"], rec["dram_bytes_per_particle"], npart))
        except Exception:
            pass

    # ---- e2e: host field buffers in, host sources out, every step
    e2e = None
    if not a.no_e2e:
        hE = torch.empty(E.t.shape, dtype=torch.float64).pin_memory()
        hB = torch.empty(B.t.shape, dtype=torch.float64).pin_memory()
        hS = torch.empty(src.t.shape, dtype=torch.float64).pin_memory()
        hE.copy_(E.t); hB.copy_(B.t)

        # Every step copies ITS E, B host -> device and ITS sources device -> host.  The
        # copies run on a second stream into double-buffered device fields, so the H2D
        # of step n+1 and the D2H of step n overlap the particle kernels of the
        # neighbouring steps (the PCIe links of an 8-GPU box give each GPU ~15 GB/s
        # when all of them copy at once: unhidden, that is a third of the step there).
        main = torch.cuda.current_stream()
        copy = torch.cuda.Stream()
        fields = [(E, B), (sk.Field(m, dtype=sk.Float3), sk.Field(m, dtype=sk.Float3))]
        srcs = [src, sk.Sources(m)]
        for f in fields[1]:
            f.boundaries_set = True
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        d2h_done = [torch.cuda.Event(), torch.cuda.Event()]
        step_done = [torch.cuda.Event(), torch.cuda.Event()]
        st = {"n": 0}

        def h2d(i):
            with torch.cuda.stream(copy):
                copy.wait_event(step_done[i])       # the last step that read this pair
                fields[i][0].t.copy_(hE, non_blocking=True)
                fields[i][1].t.copy_(hB, non_blocking=True)
                h2d_done[i].record(copy)
        for i in range(2):
            step_done[i].record(main)
            d2h_done[i].record(copy)
        h2d(0)                                      # inputs of the first step

        def step_e2e():
            i = st["n"] & 1
            st["n"] += 1
            Ei, Bi = fields[i]
            si = srcs[i]
            h2d(i ^ 1)                              # next step's inputs, behind this one's
            main.wait_event(h2d_done[i])
            main.wait_event(d2h_done[i])            # the sources buffer is free again
            ions.push(Ei, Bi, dt)
            si.deposit(ions)
            si.add_guards()
            si.copy_guards()
            step_done[i].record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(step_done[i])
                hS.copy_(si.t, non_blocking=True)
                d2h_done[i].record(copy)
        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()

        def timed_e2e(k):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for _ in range(k):
                step_e2e()
            main.wait_event(d2h_done[(st["n"] - 1) & 1])     # the last result is on the host
            e1.record(main)
            barrier()
            t = e0.elapsed_time(e1)
            return comm.allreduce(t, op=sk.comm.MAX) if size > 1 else t
        ms2 = timed_e2e(a.steps)
        e2e = {"value": n_total*a.steps/(ms2*1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(hE.numel()*8 + hB.numel()*8),
               "d2h_bytes_per_step": int(hS.numel()*8), "ms_per_step": ms2/a.steps,
               "what": "every step: E,B copied H2D from pinned host memory and that step's "
                       "sources copied D2H, on a copy stream with double-buffered device "
                       "fields (overlapping the neighbouring steps' kernels); particles "
                       "stay resident in HBM"}

    # ---- parity at the full workload size (oracle = checker), then conservation
    checks = {}
    if not a.no_parity:
        checks.update(parity_checks(sk, ions, E, B, src, dt, comm, a))
    else:
        step()
    n_now = comm.allreduce(int(ions.N), op=sk.comm.SUM) if size > 1 else int(ions.N)
    rho_sum = float(src.t[m.lby:m.uby, m.lbx:m.ubx, 0].sum().item())
    if size > 1:
        rho_sum = comm.allreduce(rho_sum, op=sk.comm.SUM)
    expect = n_total*ions.charge/a.ppc
    checks.update({"particles_conserved": n_now == n_total,
                   "charge_rel_err": abs(rho_sum - expect)/expect})

    cpu = None
    if rank == 0 and size == 1 and not a.no_cpu_baseline:
        from oracle import ref_driver
        cpu = ref_driver.measure(nx=a.nx, rows=8, ppc=a.ppc, steps=6, warmup=1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": size,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms/a.steps,
                "higher_is_better": True, "scaling": "weak" if a.weak else "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(a, {"particles_per_gpu": n_local,
                                              "layout": a.layout if not gapped else
                                              ("gapped" if accg else "dense (fallback)")}),
                "roofline": roofline, "kernels": kern, "alternative_kernels": standalone, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clk, "checks": checks, "impl": "b200"}
        print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------
CONFIGS = {
    1: dict(nx=32, ny=32, ppc=256, what="ion-acoustic wave test at its default small grid "
            "(BASELINE config 1, reference tests/test_ionacoustic.py:160-178: 45 steps per "
            "wave period): push + deposit + add_guards + copy_guards + Ohm + copy_guards "
            "per step; launch / synchronisation bound at this size"),
    2: dict(nx=1024, ny=1024, ppc=64, what="Landau / ion-acoustic loop (BASELINE config 2): "
            "push + deposit + add_guards + copy_guards + Ohm + copy_guards per step"),
    3: dict(nx=2048, ny=2048, ppc=121, what="hybrid stepper (BASELINE config 3): one Horowitz "
            "iterate per step = push_and_deposit sweep + Faraday/Hall-Ohm iterations, "
            "B = x-hat, lbx = lby = 2, quiet start 11 x 11 per cell"),
    4: dict(nx=2048, ny=2048, ppc=256, what="shearing sheet (BASELINE config 4): push_modified "
            "(S = -3/2, Omega = 1) + deposit + shear-periodic add_guards / copy_guards per "
            "step"),
}


def config_arm(a):
    """BASELINE.json configs 2-4 in the same JSON format (config 5 = b200_arm): whole-step
    timing, live roofline of the particle kernel, parity / conservation checks."""
    import numpy as np
    import torch
    import skeletor_b200 as sk
    from skeletor_b200 import _lib

    c = CONFIGS[a.config]
    if a.nx == NX and a.ny == NY and a.ppc == PPC:       # not overridden on the command line
        a.nx, a.ny, a.ppc = c["nx"], c["ny"], c["ppc"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm = sk.COMM_WORLD if world > 1 else sk.COMM_SELF
    rank, size = comm.rank, comm.size
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lb = 1 if a.config in (1, 2) else 2
    if a.config == 4:
        m = sk.ShearingManifold(a.nx, a.ny, comm, lbx=lb, lby=lb, S=-1.5, Omega=1.0,
                                Lx=1.0, Ly=a.ny/a.nx)
    elif a.config == 3:
        m = sk.Manifold(a.nx, a.ny, comm, lbx=lb, lby=lb, Lx=0.5*a.nx, Ly=0.5*a.ny)
    else:
        m = sk.Manifold(a.nx, a.ny, comm, lbx=lb, lby=lb, Lx=1.0, Ly=a.ny/a.nx)
    n_local = a.nx*m.nyp*a.ppc
    n_total = a.nx*a.ny*a.ppc
    # slot ranges of the gapped layout: 25 % slack per cell + 4 slots for every empty cell
    # of the 16 x 16 tiles that cover the extended grid
    nmax = int(1.36*n_local) + 4096 + 4*256*((m.mx + 15)//16)*((m.myp + 15)//16)
    nmax += nmax & 1
    ions = sk.Particles(m, nmax, charge=1.0, mass=1.0, order=1,
                        nbmax=max(n_local//100, 1 << 16))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    d = ions._data
    if a.config == 3:
        # quiet start: sq x sq sub-lattice per cell, cold-ish ions (a noisy start drives
        # O(1) electric fields through grad ln(rho) in the Hall / pressure terms)
        sq = int(round(a.ppc**0.5))
        assert sq*sq == a.ppc, "config 3 needs a square number of particles per cell"
        ax = (torch.arange(a.nx*sq, device=dev, dtype=torch.float64) + 0.5)/sq
        ay = m.noff + (torch.arange(m.nyp*sq, device=dev, dtype=torch.float64) + 0.5)/sq
        d[0, :n_local] = ax.repeat(m.nyp*sq)
        d[1, :n_local] = ay.repeat_interleave(a.nx*sq)
        d[2:5, :n_local] = 0.1*torch.randn((3, n_local), generator=gen, device=dev,
                                           dtype=torch.float64)
    else:
        d[0, :n_local] = torch.rand(n_local, generator=gen, device=dev,
                                    dtype=torch.float64)*a.nx
        d[1, :n_local] = m.noff + torch.rand(n_local, generator=gen, device=dev,
                                             dtype=torch.float64)*m.nyp
        vt = 0.05 if a.config == 4 else 1.0
        d[2:5, :n_local] = vt*torch.randn((3, n_local), generator=gen, device=dev,
                                          dtype=torch.float64)
    ions.N = n_local
    del d               # (no reference to the tensor: Particles swaps / releases its buffers)
    E = sk.Field(m, dtype=sk.Float3)
    B = sk.Field(m, dtype=sk.Float3)
    src = sk.Sources(m)
    hot = ["skb_push_gapped", "skb_boris_push", "skb_deposit", "skb_push_and_deposit_gapped",
           "skb_push_and_deposit"]
    stepper = None
    if a.config in (1, 2):
        dt = 0.1*m.dx
        E.copy_guards(); B.copy_guards()
        ohm = sk.Ohm(m, temperature=1.0, charge=1.0)
        src.deposit(ions, set_boundaries=True)

        def step():
            ions.push(E, B, dt)
            src.deposit(ions)
            src.add_guards()
            src.copy_guards()
            ohm(src, B, E)
            E.copy_guards()
        sweep_bytes, dom_ep = 120.0, "skb_push_gapped"
    elif a.config == 4:
        dt = 0.1*m.dx
        xg, yg = np.meshgrid(m.x, m.y)
        E['x'].active = 0.01*np.sin(2*np.pi*xg/m.Lx)
        B['z'].active = 1.0
        E.copy_guards(); B.copy_guards()

        def step():
            ions.push_modified(E, B, dt)
            src.deposit(ions)
            src.time = ions.time
            src.add_guards()
            src.copy_guards()
        sweep_bytes, dom_ep = 120.0, "skb_push_gapped"
    else:
        from skeletor_b200.time_steppers.horowitz import TimeStepper as Horowitz
        dt = 1e-2
        B.fill((1.0, 0.0, 0.0))
        B.copy_guards()
        stepper = Horowitz(sk.State(ions, B), sk.Ohm(m, temperature=0.01, charge=1.0), m)
        # (no prepare(): the iteration to t = 0 consistency is set-up, not the step)
        ions.deposit(set_boundaries=True)
        stepper.sources.t.copy_(ions.sources.t)
        stepper.sources.boundaries_set = True
        stepper.ohm(stepper.sources, stepper.B, stepper.E, set_boundaries=True)

        def step():
            stepper.iterate(dt)
        sweep_bytes, dom_ep = 80.0, "skb_push_and_deposit_gapped"

    def barrier():
        if size > 1:
            comm.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if size > 1:
            ms = comm.allreduce(ms, op=sk.comm.MAX)
        return ms

    for _ in range(max(a.warmup, 3)):
        step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    k0 = _lib.kernel_launches
    _lib.trace = {k: [] for k in hot}
    ms = timed(step, a.steps)
    live = {k: [e0.elapsed_time(e1) for e0, e1 in v] for k, v in _lib.trace.items() if v}
    _lib.trace = None
    launches = _lib.kernel_launches - k0
    clk = clocks.stop() if rank == 0 else None
    value = n_total*a.steps/(ms*1e-3)
    peak, peak_src = measured_peak()
    npart = int(ions.N)
    cells = m.mx*m.myp
    alg = {"skb_push_gapped": 80.0*npart + 48.0*cells, "skb_boris_push": 80.0*npart + 48.0*cells,
           "skb_deposit": 40.0*npart + 32.0*cells,
           "skb_push_and_deposit_gapped": 80.0*npart + 80.0*cells,
           "skb_push_and_deposit": 80.0*npart + 80.0*cells}
    kern = {}
    for ep, ts in live.items():
        ts = [t for t in ts if t > 0.2*max(ts)]
        avg = sum(ts)/len(ts)
        kern[ep] = {"live_ms": round(avg, 4), "live_launches": len(ts),
                    "launches_per_step": round(len(ts)/a.steps, 2),
                    "alg_bytes": alg[ep],
                    "live_frac": round(alg[ep]/(avg*1e-3)/1e9/peak, 4)}
    dom = dom_ep if dom_ep in kern else max(kern, key=lambda k: kern[k]["live_ms"])
    dk = kern[dom]
    roofline = {"kernel": dom, "bound": "hbm",
                "achieved": round(dk["alg_bytes"]/(dk["live_ms"]*1e-3)/1e9, 1), "peak": peak,
                "unit": "GB/s", "frac": dk["live_frac"], "traffic": None,
                "peak_source": peak_src, "alg_bytes_per_launch": dk["alg_bytes"],
                "ms_per_launch": dk["live_ms"],
                "timing": "CUDA events around the launch in every timed step, average",
                "step_bytes_per_particle": sweep_bytes,
                "step_frac": round(sweep_bytes*npart/(ms/a.steps*1e-3)/1e9/peak, 4),
                "others": {k: v for k, v in kern.items() if k != dom}}
    layout = ions._rep
    checks = {}
    if a.config in (1, 2, 4) and not a.no_parity:
        checks.update(parity_checks(sk, ions, E, B, src, dt, comm, a,
                                    modified=a.config == 4))
    n_now = comm.allreduce(int(ions.N), op=sk.comm.SUM) if size > 1 else int(ions.N)
    checks["particles_conserved"] = n_now == n_total
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": size,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms/a.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": c["what"], "baseline_config": a.config,
                           "grid": [a.nx, a.ny], "ppc": a.ppc, "particles": n_total,
                           "interpolation": "CIC", "decomposition": "y-slabs, 1 per GPU",
                           "particles_per_gpu": n_local, "layout": layout,
                           "l2_policy": "inputs larger than L2"},
                "roofline": roofline, "kernels": kern, "cpu_baseline": None, "e2e": None,
                "gpu_launches": launches, "clocks": clk, "checks": checks, "impl": "b200"}
        if a.config == 1:
            # the reference's whole test file (import + 45 steps, quiet start) takes 22.5 s
            # single-rank on the survey's CPU
            line["config"]["seconds_per_45_steps"] = 45*ms/a.steps*1e-3
        print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.weak:
        a.ny = a.ny*max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if a.impl == "reference":
        reference_arm(a)
    elif a.config != 5:
        config_arm(a)
    else:
        b200_arm(a)


if __name__ == "__main__":
    main()
