#!/usr/bin/env python
"""bench.py — benchmark of the RHS hot path (flux_div + ghost exchange + RK stage update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]

Workloads (BASELINE.json `configs`, numbered from 1):
  2 (default, the configuration the metric is quoted on): Taylor-Green vortex, 512^3 cells per GPU in 32^3 blocks
    (16 x 16 x 16N blocks, 2 exchange cells), totani_lr + visc_lr, rk4_t with the fused prim/cons update, periodic.
  4: the weak-scaling sweep, 256^3 cells per GPU (8 x 8 x 8N blocks), same solver. The default run also measures
    configs 4, 3 and 5 briefly and reports them in the `configs` sub-records of the same JSON line.
  3: compressible channel, 1024 x 512 x 64N cells (32 x 16 x 2N blocks; N = 8 is the 1024 x 512 x 512 grid), y stretched
    with integrated_tanh_1D, isothermal no-slip walls, hybrid(totani_lr, fweno_t, ducros_t) + visc_lr. FP64-bound:
    `roofline.bound` = "fp64".
  5: AMR block grid (8^3 roots of 32^3 cells, two sphere refinements, 1912 blocks = 62.7 M cells) partitioned over N GPUs
    with SPADE's contiguous partition; block boxes and transaction tables are SPADE's own (tests/golden/config5_amr.npz,
    written by tests/golden/make_config5.py from the unmodified reference). Fixed grid: "scaling": "strong".
One "step" = one RK4 time step = 4 x (flux_div + exchange + stage update) over the whole grid.

metric: cell-stage-updates/s = cells x stages x steps / time (a cell advanced through one RK stage), whole job.
Before the timed region every run advances a small lattice of the same functor set through the same code path
(fused stage kernel, two-stream overlap, peer-memory or NCCL messages) and compares it with the oracle
(`parity_check` in the line: exchange bit-exact, 2 RK4 steps to 1e-12). Prints ONE JSON line.

Development switches (environment; none of them changes what the default line measures): SPB_PHASE_EVENTS=1 adds per-rank
stage phases, join / step times, clocks and GPU identity (`phases`); SPB_DEFER_UNPACK=0, SPB_P2P=0, SPB_BLOCK_RUNS=1 select the
undeferred schedule, NCCL send/recv, one launch per contiguous block run; SPB_DIAG_SOLO=1 lets every rank run the 1-GPU
workload under torchrun; SPB_DIAG_DELAY_RANK=r delays rank r by 30 ms at the head of the timed region; SPB_BOUNDARY_DELAY_US
delays the boundary kernel on the side stream; SPB_NO_CLOCK_SAMPLER=1 switches the NVML sampler off (tools/gpu_visit.sh phases /
solo2 / steps use them; the findings are in DESIGN section 6).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMMA, RGAS, P0, T0, U0 = 1.4, 287.15, 101325.0, 300.0, 34.7
REYNOLDS, PRANDTL = 1600.0, 0.72
STAGES = 4
BLOCK = 32          # cells per block edge
NG = 2
LATTICE = {2: (16, 16, 16), 4: (8, 8, 8), 3: (32, 16, 2)}       # blocks per GPU
# algorithmic flop per cell of the hybrid WENO + central + viscous RHS and the FP64 roofs it is held against:
# NVIDIA's B200 figure, and the DFMA issue rate measured on this pool (tools/probes/pipes_probe.cu, profiles/r02_pipes_probe.log:
# 1.45 warp-DFMA per SM per clock x 64 flop x 148 SMs x 1.965 GHz)
HYBRID_FLOP_PER_CELL = 2000.0
FP64_PEAK_NOMINAL_TFLOPS = 37.0
FP64_PEAK_MEASURED_TFLOPS = 1.45 * 64 * 148 * 1.965e9 / 1e12
METRIC = "fp64 cell-updates/sec (RHS+exchange+RK)"
UNIT = "cell-stage-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--lattice", type=int, nargs=3, default=None, help="development: blocks per GPU instead of the config's lattice")
    ap.add_argument("--scheme", default=None, choices=["central", "hybrid"], help="development: functor set instead of the config's")
    ap.add_argument("--unfused", action="store_true", help="two kernels per stage (flux_div, then rk_update) instead of the fused stage kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config-4 sub-record of the default run")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    if a.scheme is None:
        a.scheme = "hybrid" if a.config == 3 else "central"
    return a


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per interior cell of one launch, from the tracked summary of the
    `ncu --set full` captures (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic). None if no capture
    of that kernel is tracked."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(p))[kernel_key]
        return float(rec["dram_bytes_per_cell"]), rec.get("source")
    except Exception:
        return None, None


class ClockSampler:
    """SM clock / throttle reasons during the timed region (nvidia-smi, B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz, self.power = index, [], set(), None, []
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._thr = None

    def _run_nvml(self):
        """NVML directly (nvidia_ml_py): a query takes well under a millisecond, so even a 0.2 s timed region gets
        tens of samples; same quantities as the nvidia-smi line of the recipe."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self._ready.set()
        while not self._stop.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for n, b in bits.items():
                if r & b:
                    self.reasons.add(n)
            self._stop.wait(0.005)

    def _run(self):
        try:
            return self._run_nvml()
        except Exception:
            pass                      # no NVML binding: the nvidia-smi line of the recipe
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self._ready.set()
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        """Returns once the sampler is initialised (module import, nvmlInit, device handle): that costs ~10 ms of driver and
        interpreter time, which must not fall into the timed region that starts right after."""
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        self._ready.wait(timeout=10)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)
        med = statistics.median(self.samples) if self.samples else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.power:
            out["power_w"] = statistics.median(self.power)
        return out


# ------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own implementation of the path (oracle/_ref = unmodified SPADE headers compiled in the dev
# container with g++ -O3; falls back to the C port) on the host cores, on a bounded sample of the workload
def cpu_sample(args, steps, all_threads=True):
    import numpy as np
    from oracle import ref, port
    cores = os.cpu_count() or 1
    kind = "reference" if ref.available() else "port"
    flags = ref.use_timing_build() if kind == "reference" else "gcc -std=c11 -O2 -ffp-contract=off (oracle/spade_oracle.c)"
    lat = (4, 4, 4)                      # 128^3 cells in 32^3 blocks: the block shape of the GPU workload
    nranks = max(1, min(cores, lat[0] * lat[1] * lat[2])) if (kind == "reference" and all_threads) else 1
    scheme = 1 if args.scheme == "hybrid" else 0
    mu = (P0 / (RGAS * T0)) * U0 * 1.0 / REYNOLDS
    cfg = ref.make_cfg(lat, (BLOCK,) * 3, NG, scheme=scheme, gamma=GAMMA, R=RGAS, mu=mu, prandtl=PRANDTL,
                       sensor_eps=1e-2, nranks=nranks, integrator=0)
    q = host_state(lat, np)
    dt = 0.2 * (2 * np.pi / (lat[0] * BLOCK)) / port.reduce_umax(cfg, q.ravel())
    cells = (lat[0] * BLOCK) ** 3
    if kind == "reference":
        qq, _ = ref.advance(cfg, q.ravel(), dt, 1)              # warm-up step (first touch of the thread pools)
        qq, sec = ref.advance(cfg, qq, dt, steps)
    else:
        qq = port.advance(cfg, q.ravel(), dt, 1)
        t0 = time.time()
        port.advance(cfg, qq, dt, steps)
        sec = time.time() - t0
    name = "hybrid(totani_lr,fweno_t,ducros_t) + visc_lr" if scheme else "totani_lr + visc_lr"
    sample = (f"TGV {lat[0]*BLOCK}^3 cells in {BLOCK}^3 blocks ({lat[0]}x{lat[1]}x{lat[2]}), {name}, rk4, {steps} steps in {sec:.1f} s: a bounded "
              f"sample of the workload (same block shape, functor set, integrator; the CPU path's cost per cell does not depend on the grid size)")
    return {"value": cells * STAGES * steps / sec, "unit": UNIT, "cores": nranks, "kind": kind, "sample": sample,
            "flags": flags, "seconds": sec, "ms_per_step": 1e3 * sec / steps}


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cb = cpu_sample(args, max(1, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.config == 5 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus, sample_note=cb["sample"]),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "flags")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def host_state(lat, np, block=(BLOCK, BLOCK, BLOCK), perturb=0.0, seed=0, bounds_hi=None):
    """TGV initial condition on the host (all cells incl. ghosts analytic), reference memory order."""
    nlb = lat[0] * lat[1] * lat[2]
    L = 2 * np.pi
    hi = bounds_hi or (L, L, L)
    q = np.zeros((nlb, block[2] + 2 * NG, block[1] + 2 * NG, block[0] + 2 * NG, 5))
    rho0 = P0 / (RGAS * T0)
    for lb in range(nlb):
        b = (lb % lat[0], (lb // lat[0]) % lat[1], lb // (lat[0] * lat[1]))
        ax = [b[d] * hi[d] / lat[d] + (np.arange(-NG, block[d] + NG) + 0.5) * (hi[d] / lat[d] / block[d]) for d in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        q[lb, ..., 0] = P0 + rho0 * U0 * U0 / 16 * (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2)
        q[lb, ..., 1] = T0 * (1 + (0.02 * np.sin(X + 2 * Y - Z) if perturb else 0.0))
        q[lb, ..., 2] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
        q[lb, ..., 3] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
        q[lb, ..., 4] = (0.3 * U0 * np.sin(Z) * np.cos(X + Y)) if perturb else 0.0
    if perturb:
        q *= 1 + perturb * np.random.default_rng(12345 + seed).uniform(-1, 1, q.shape)
    return q


def workload_config(args, n, sample_note=None):
    c = args.config
    lat = tuple(args.lattice) if args.lattice else LATTICE.get(c)
    sch = "totani_lr" if args.scheme == "central" else "hybrid(totani_lr,fweno_t,ducros_t(1e-2),full_flux)"
    if c in (2, 4):
        cfg = {"workload": f"BASELINE config {c}: TGV {lat[0]*BLOCK}x{lat[1]*BLOCK}x{lat[2]*BLOCK*n} cells ({lat[0]}x{lat[1]}x{lat[2]*n} blocks of {BLOCK}^3, "
                           f"{NG} exchange cells), {sch} + visc_lr, rk4_t fused prim/cons, periodic",
               "cells_per_gpu": lat[0] * lat[1] * lat[2] * BLOCK ** 3,
               "partition": f"contiguous block runs, rank r = z-slab r ({n} ranks)",
               "l2": "inputs larger than L2 (q + scratch + 4 residual arrays, %.1f GB each)" % (lat[0] * lat[1] * lat[2] * (BLOCK + 2 * NG) ** 3 * 40 / 1e9)}
    elif c == 3:
        cfg = {"workload": f"BASELINE config 3: channel {lat[0]*BLOCK}x{lat[1]*BLOCK}x{lat[2]*BLOCK*n} cells ({lat[0]}x{lat[1]}x{lat[2]*n} blocks of {BLOCK}^3), "
                           f"y = integrated_tanh_1D(-1, 1, 0.1, 1.3), {sch} + visc_lr, rk4_t, no-slip isothermal walls in y, x/z periodic"
                           + (" (N = 8 is the 1024x512x512 grid)" if n != 8 else ""),
               "cells_per_gpu": lat[0] * lat[1] * lat[2] * BLOCK ** 3,
               "partition": f"contiguous block runs, rank r = z-slab r ({n} ranks)",
               "l2": "inputs larger than L2 (6 arrays of 1.9 GB per GPU)"}
    else:
        cfg = {"workload": "BASELINE config 5: AMR block grid, 8x8x8 roots of 32^3 cells on [0,2pi)^3, blocks intersecting |x-c| < 0.30*2pi refined, "
                           "their children intersecting |x-c| < 0.12*2pi refined again (amr::constraints::factor2): 1912 blocks = 62.7 M cells, 3 levels; "
                           f"{sch} + visc_lr, rk4_t, periodic; block boxes and exchange tables from the reference (tests/golden/config5_amr.npz)",
               "cells_total": 1912 * BLOCK ** 3,
               "partition": f"spade::partition contiguous runs of global block ids over {n} ranks (every block costs the same)",
               "l2": "inputs larger than L2 (6 arrays of 3.6 GB in total)"}
    cfg["stages_per_step"] = STAGES
    if sample_note:
        cfg["reference_sample"] = sample_note
    return cfg


# ------------------------------------------------------------------------------------------------------
def tgv_device_state(sp, grid, torch):
    """TGV initial condition generated on the device block-batch by block-batch (torch is plumbing here)."""
    nlb = grid.num_local_blocks
    q = sp.grid_array(grid, 0.0, (NG,) * 3)
    rho0 = P0 / (RGAS * T0)
    idx = torch.arange(-NG, BLOCK + NG, dtype=torch.float64, device="cuda") + 0.5
    chunk = 256
    for b0 in range(0, nlb, chunk):
        b1 = min(nlb, b0 + chunk)
        org = torch.tensor([grid.blocks.get_block_box(grid.first_block + l)[0::2] for l in range(b0, b1)],
                           dtype=torch.float64, device="cuda")
        dx = [grid.get_dx(d) for d in range(3)]
        X = (org[:, 0, None] + idx[None, :] * dx[0])[:, None, None, :]
        Y = (org[:, 1, None] + idx[None, :] * dx[1])[:, None, :, None]
        Z = (org[:, 2, None] + idx[None, :] * dx[2])[:, :, None, None]
        v = q.data[b0:b1]
        v[..., 0] = P0 + rho0 * U0 * U0 / 16 * (torch.cos(2 * X) + torch.cos(2 * Y)) * (torch.cos(2 * Z) + 2)
        v[..., 1] = T0
        v[..., 2] = U0 * torch.sin(X) * torch.cos(Y) * torch.cos(Z)
        v[..., 3] = -U0 * torch.cos(X) * torch.sin(Y) * torch.cos(Z)
        v[..., 4] = 0.0
    return q


device_state = tgv_device_state          # tools/kbench.py


def boxes_device_state(sp, grid, boxes, torch):
    """the same TGV field on blocks given by their boxes (AMR: per-block spacing)"""
    q = sp.grid_array(grid, 0.0, (NG,) * 3)
    rho0 = P0 / (RGAS * T0)
    idx = torch.arange(-NG, BLOCK + NG, dtype=torch.float64, device="cuda") + 0.5
    bx = torch.tensor(boxes, dtype=torch.float64, device="cuda")
    for b0 in range(0, grid.num_local_blocks, 256):
        b1 = min(grid.num_local_blocks, b0 + 256)
        lo = bx[b0:b1, 0::2]
        dx = (bx[b0:b1, 1::2] - lo) / BLOCK
        X = (lo[:, 0, None] + idx[None, :] * dx[:, 0, None])[:, None, None, :]
        Y = (lo[:, 1, None] + idx[None, :] * dx[:, 1, None])[:, None, :, None]
        Z = (lo[:, 2, None] + idx[None, :] * dx[:, 2, None])[:, :, None, None]
        v = q.data[b0:b1]
        v[..., 0] = P0 + rho0 * U0 * U0 / 16 * (torch.cos(2 * X) + torch.cos(2 * Y)) * (torch.cos(2 * Z) + 2)
        v[..., 1] = T0
        v[..., 2] = U0 * torch.sin(X) * torch.cos(Y) * torch.cos(Z)
        v[..., 3] = -U0 * torch.cos(X) * torch.sin(Y) * torch.cos(Z)
        v[..., 4] = 0.0
    return q


def contiguous_partition(nglob, nranks, rank):
    """spade::partition::block_partition_t (grid/partition.h:27-84): contiguous runs, the first nglob % nranks ranks one extra"""
    per, extra = divmod(nglob, nranks)
    first = rank * per + min(rank, extra)
    return first, per + (1 if rank < extra else 0)


def amr_tables(fix, prefix, nranks, rank, np):
    return tuple(np.ascontiguousarray(fix[f"{prefix}_{k}_{nranks}_{rank}"], dtype=np.int64) for k in ("send", "recv", "isend", "irecv"))


class Workload:
    pass


def make_flux(sp, gas, scheme, mu):
    conv = sp.totani_lr(gas)
    if scheme == "hybrid":
        conv = sp.hybrid_scheme_t(conv, sp.fweno_t(gas), sp.ducros_t(1e-2), sp.full_flux)
    return sp.flux_desc(sp.compose(conv, sp.visc_lr(sp.constant_viscosity_t(mu, PRANDTL), gas)))


def build_workload(cfgid, args, sp, pool, torch, timing):
    import numpy as np
    w = Workload()
    n = pool.size()
    w.cfgid, w.n = cfgid, n
    w.gas = gas = sp.ideal_gas_t(GAMMA, RGAS)
    w.scheme = args.scheme if cfgid == args.config else ("hybrid" if cfgid == 3 else "central")
    w.scaling = "strong" if cfgid == 5 else "weak"
    bc_walls = None
    if cfgid in (2, 4):
        lat = tuple(args.lattice) if (args.lattice and cfgid == args.config) else LATTICE[cfgid]
        L = 2 * np.pi
        blocks = sp.cartesian_blocks_t((lat[0], lat[1], lat[2] * n), [0.0, L, 0.0, L, 0.0, L * n])
        w.grid = grid = sp.cartesian_grid_t((BLOCK,) * 3, blocks, sp.identity(), pool)
        mu = (P0 / (RGAS * T0)) * U0 * 1.0 / REYNOLDS
        w.q = tgv_device_state(sp, grid, torch)
        w.handle = sp.make_exchange(w.q, (True, True, True))
        dmin = grid.get_dx(0)
    elif cfgid == 3:
        lat = tuple(args.lattice) if (args.lattice and cfgid == args.config) else LATTICE[3]
        pi = float(np.pi)
        u0 = 69.4
        blocks = sp.cartesian_blocks_t((lat[0], lat[1], lat[2] * n), [0.0, 4 * pi, -1.0, 1.0, 0.0, 2 * pi * n / 8.0])
        coords = sp.diagonal_coords(None, sp.integrated_tanh_1D(-1.0, 1.0, 0.1, 1.3), None)
        w.grid = grid = sp.cartesian_grid_t((BLOCK,) * 3, blocks, coords, pool)
        mu = (P0 / (RGAS * T0)) * u0 / 3000.0
        # laminar parabolic profile + seeded perturbation + a planted pressure jump that wakes the shock sensor
        w.q = q = sp.grid_array(grid, 0.0, (NG,) * 3)
        idx = torch.arange(-NG, BLOCK + NG, dtype=torch.float64, device="cuda") + 0.5
        gen = torch.Generator(device="cuda").manual_seed(12345 + pool.rank())
        for b0 in range(0, grid.num_local_blocks, 128):
            b1 = min(grid.num_local_blocks, b0 + 128)
            org = torch.tensor([blocks.get_block_box(grid.first_block + l)[0::2] for l in range(b0, b1)], dtype=torch.float64, device="cuda")
            dx = [grid.get_dx(d) for d in range(3)]
            X = (org[:, 0, None] + idx[None, :] * dx[0])[:, None, None, :]
            Y = (org[:, 1, None] + idx[None, :] * dx[1])[:, None, :, None]
            Z = (org[:, 2, None] + idx[None, :] * dx[2])[:, :, None, None]
            v = q.data[b0:b1]
            v[..., 0] = P0 * torch.where(torch.sin(0.5 * X) > 0.3, 1.2, 1.0) + 0 * Y + 0 * Z
            v[..., 1] = T0 + 0 * X + 0 * Y + 0 * Z
            v[..., 2] = u0 * (1 - Y * Y) + 0 * X + 0 * Z
            v[..., 3] = 0.02 * u0 * torch.sin(X) * torch.cos(pi * Y) * torch.cos(Z)
            v[..., 4] = 0.02 * u0 * torch.sin(Z) * torch.cos(X + pi * Y)
            v *= 1 + 1e-3 * (2 * torch.rand(v.shape, dtype=torch.float64, device="cuda", generator=gen) - 1)
        w.handle = sp.make_exchange(q, (True, False, True))
        bc_walls = (sp.boundary.ymin | sp.boundary.ymax, sp.noslip_isothermal_wall(T0))
        _, jac, _ = grid.metric_tables((NG,) * 3)
        dmin = min(grid.get_dx(0), grid.get_dx(2), float(np.abs(jac[1][:, NG:-NG]).min()) * grid.get_dx(1))
    else:
        fix = np.load(os.path.join(ROOT, "tests", "golden", "config5_amr.npz"))
        boxes = fix["b_boxes"]
        if n not in (1, 2, 4, 8):
            raise SystemExit("bench.py --config 5: the tracked tables cover 1, 2, 4 and 8 ranks")
        first, cnt = contiguous_partition(len(boxes), n, pool.rank())
        w.grid = grid = sp.cartesian_grid_t.from_boxes((BLOCK,) * 3, boxes[first:first + cnt], pool, first_block=first)
        mu = (P0 / (RGAS * T0)) * U0 * 1.0 / REYNOLDS
        w.q = boxes_device_state(sp, grid, boxes[first:first + cnt], torch)
        w.handle = sp.make_exchange(w.q, (True, True, True), tables=amr_tables(fix, "b", n, pool.rank(), np))
        dmin = float((boxes[:, 1] - boxes[:, 0]).min()) / BLOCK
        w.cells_total = len(boxes) * BLOCK ** 3
    w.flux = make_flux(sp, gas, w.scheme, mu)
    w.rhs = sp.grid_array(w.grid, 0.0, (NG,) * 3)
    w.bc = sp.exchange_bc_t(w.handle, *(bc_walls or ()))
    w.bc(w.q, 0.0)
    w.umax0 = sp.transform_reduce(w.q, sp.FN_WAVESPEED, sp.RED_MAX, gas)
    w.dt = 0.2 * dmin / w.umax0
    w.fused = not args.unfused
    if w.fused:
        calc_rhs = sp.flux_div_rhs_t(w.flux, sp.overwrite)
    else:
        def calc_rhs(r, qq, t):
            if timing["on"]:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sp.flux_div(qq, r, w.flux, sp.overwrite)
                e1.record()
                timing["events"].append((e0, e1, 80.0))
            else:
                sp.flux_div(qq, r, w.flux, sp.overwrite)
    alg = sp.rk4_t
    # the usual SPADE solver set-up: integrator_t(axis, alg, data, rhs callback, boundary callback, state transform). As a
    # flux_div_rhs_t / exchange_bc_t the callbacks are recognised and every stage is ONE kernel per block range, the ghost
    # messages of the rank-boundary blocks fly while the rank-interior blocks are advanced
    w.ti = sp.integrator_t(sp.time_axis_t(0.0, w.dt), alg, sp.integrator_data_t(w.q, w.rhs, alg), calc_rhs, w.bc, sp.state_transform_t(gas))
    w.cells_local = w.grid.local_cells()
    if cfgid != 5:
        w.cells_total = w.cells_local * n
    return w


def measure(w, steps, warmup, sp, torch, dist, world, rank, local_rank, timing, sample_clocks):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        w.ti.advance()
    # the byte count of a stage launch needs the plan's same-rank ghost cells: read from the tables on first use (~15 ms of host time
    # for 4096 blocks, cached afterwards) — before the timed region, not inside its first step (rounds 1 and 2 paid it there:
    # first step 28.7 ms against a median of 19.4 ms, `steps_detail` in the line)
    if getattr(w, "handle", None) is not None and hasattr(w.handle, "local_injection_cells"):
        w.handle.local_injection_cells()
    barrier()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(vis.split(",")[local_rank]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local_rank   # NVML counts physical GPUs
    sampler = ClockSampler(phys)
    sample_clocks = sample_clocks and not os.environ.get("SPB_NO_CLOCK_SAMPLER")     # development A/B only
    all_sample = bool(os.environ.get("SPB_PHASE_EVENTS"))                            # diagnosis: every rank samples its own GPU
    if (rank == 0 or all_sample) and sample_clocks:
        sampler.start()
    launches0 = sp.launch_count()
    timing["events"] = ev = []
    timing["on"] = True
    if os.environ.get("SPB_PHASE_EVENTS"):
        w.handle.finish_trace = []
    if w.fused:
        w.ti.stage_events = ev
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if os.environ.get("SPB_DIAG_DELAY_RANK") == str(rank):
        time.sleep(0.03)                     # diagnosis: this rank enters the timed region late (who runs ahead of whom?)
    t0.record()
    host_t0 = time.perf_counter()
    step_ev, host_ms = [], []
    for _ in range(steps):
        h0 = time.perf_counter()
        w.ti.advance()
        host_ms.append(1e3 * (time.perf_counter() - h0))
        step_ev.append(torch.cuda.Event(enable_timing=True))
        step_ev[-1].record()
    host_loop_ms = 1e3 * (time.perf_counter() - host_t0)      # diagnosis: far below the device time if the host runs ahead
    t1.record()
    barrier()
    timing["on"] = False
    w.ti.stage_events = None
    ms = t0.elapsed_time(t1)
    launches = sp.launch_count() - launches0
    clocks = sampler.stop() if ((rank == 0 or all_sample) and sample_clocks) else None
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev) / max(1, len(ev))
    kern_bpc = sum(c for _, _, c in ev) / max(1, len(ev))      # algorithmic bytes per cell, mean over launches
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    phases = None
    if os.environ.get("SPB_PHASE_EVENTS"):
        # diagnosis (SPB_PHASE_EVENTS=1), per rank: stage kernel time, own step time, clocks of the own GPU; with the overlapped multi-rank
        # schedule also the mean device times of a stage: start -> boundary kernel done -> packs and flags issued -> interior kernel
        # done -> messages unpacked, the end-of-step join and the [flag wait, unpack] pairs
        props = torch.cuda.get_device_properties(torch.cuda.current_device())
        phases = {"stage_kernel_ms": sum(x.elapsed_time(y) for x, y, _ in ev) / max(1, len(ev)), "ms_per_step_own": t0.elapsed_time(t1) / steps,
                  "host_loop_ms_per_step": host_loop_ms / steps, "clocks": clocks,
                  "gpu": {"local_rank": local_rank, "nvml_index": phys, "uuid": str(getattr(props, "uuid", ""))[-8:]}}
        pes = (getattr(w.ti, "phase_events", None) or [])[-4 * steps:]
        if pes:
            mean = lambda f: sum(f(p) for p in pes) / len(pes)
            phases.update({"boundary_done_ms": mean(lambda p: p[0].elapsed_time(p[1])), "packs_done_ms": mean(lambda p: p[0].elapsed_time(p[2])),
                           "interior_done_ms": mean(lambda p: p[0].elapsed_time(p[3])), "unpacked_ms": mean(lambda p: p[0].elapsed_time(p[4])), "stages": len(pes)})
            w.ti.phase_events.clear()
        je = getattr(w.ti, "join_events", None)
        if je:
            je = je[-steps:]
            phases["join_ms"] = sum(x.elapsed_time(y) for x, y in je) / len(je)                 # end-of-step join (deferred schedule)
            phases["step_ms"] = sum(je[i][1].elapsed_time(je[i + 1][1]) for i in range(len(je) - 1)) / max(1, len(je) - 1)
            w.ti.join_events.clear()
        tr = getattr(getattr(w, "handle", None), "finish_trace", None)
        if tr:
            tr = tr[-4 * steps:]
            phases["finish_ms"] = [round(sum(t[k].elapsed_time(t[k + 1]) for t in tr) / len(tr), 4) for k in range(len(tr[0]) - 1)]   # wait, unpack per peer
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, phases)
            phases = {"per_rank": allp}
    # device time of every step (events behind each advance()) and host time of every advance() call: the first step of the region
    # starts on an idle GPU with an empty launch queue, the others are enqueued while their predecessor runs
    per_step = [t0.elapsed_time(step_ev[0])] + [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(steps - 1)]
    steps_detail = {"first_ms": per_step[0], "median_ms": statistics.median(per_step), "max_ms": max(per_step),
                    "host_first_ms": host_ms[0], "host_median_ms": statistics.median(host_ms)}
    umax_end = sp.transform_reduce(w.q, sp.FN_WAVESPEED, sp.RED_MAX, w.gas)
    if not (umax_end == umax_end) or umax_end > 10 * w.umax0:
        raise SystemExit(f"bench.py: solution diverged (umax {umax_end})")
    return {"ms": ms, "launches": int(launches), "clocks": clocks, "phases": phases, "kern_ms": kern_ms, "kern_bpc": kern_bpc, "n_events": len(ev),
            "steps_detail": steps_detail,
            "value": w.cells_total * STAGES * steps / (ms * 1e-3)}


def roofline_of(w, m, steps):
    """Roofline of the dominant kernel of one rank. Algorithmic bytes per interior cell: plain flux_div
    reads q and writes rhs = 80 B; the fused stage kernel reads q, writes q', reads/writes the residual registers its stage
    needs (rk4 = 120, 160, 200, 120 B, mean 150 B per launch) and writes the same-rank ghost cells (40 B x 0.4238 ghost
    cells per interior cell at n = 32, g = 2 = 17 B). The hybrid WENO set is FP64-bound (2 000 algorithmic flop per cell)."""
    peak, peak_src = measured_peak_hbm()
    hbm_achieved = m["kern_bpc"] * w.cells_local / (m["kern_ms"] * 1e-3) / 1e9
    common = {"alg_bytes_per_cell": m["kern_bpc"], "ms_per_launch": m["kern_ms"], "launches_timed": m["n_events"],
              "cell_evals_per_s": w.cells_local / (m["kern_ms"] * 1e-3), "step_share": m["kern_ms"] * STAGES * steps / m["ms"],
              "timed_with": "CUDA events around every stage launch on the launching stream, inside the timed region"}
    if w.scheme == "hybrid":
        tf = HYBRID_FLOP_PER_CELL * w.cells_local / (m["kern_ms"] * 1e-3) / 1e12
        r = {"bound": "fp64", "kernel": "flux_div_kernel<hybrid WENO + viscous, FUSED stage>" if w.fused else "flux_div_kernel<hybrid> (rhs only)",
             "achieved": tf, "peak": FP64_PEAK_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_PEAK_NOMINAL_TFLOPS,
             "peak_source": "nominal B200 FP64 (no fp64 entry in MEASURED_PEAKS.json)", "alg_flop_per_cell": HYBRID_FLOP_PER_CELL,
             "frac_of_measured_dfma_rate": tf / FP64_PEAK_MEASURED_TFLOPS, "measured_dfma_peak_tflops": FP64_PEAK_MEASURED_TFLOPS,
             "hbm": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src}}
        key = "hybrid_stage"
    else:
        r = {"bound": "hbm", "kernel": "flux_div_narrow_kernel<FUSED stage>" if w.fused else "flux_div kernel (rhs only)",
             "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src}
        key = ("fused_stage" if w.fused else "rhs") + ("_amr" if w.cfgid == 5 else "")
    tpc, tsrc = measured_traffic(key)
    r["traffic"] = tpc * w.cells_local if tpc else None
    if tsrc:
        r["traffic_source"] = tsrc
    r.update(common)
    return r


# ------------------------------------------------------------------------------------------------------
def parity_check(cfgid, scheme, sp, pool, torch):
    """A small grid of the same functor set advanced through the SAME code path as the timed run (fused stage kernel with the
    ghost warp / owner stores, rank-boundary blocks first on a side stream, peer-memory or NCCL messages), against the oracle
    (oracle.port: the C restatement pinned against the unmodified reference; used here as the checker only). Every rank holds
    the whole oracle result and compares its own blocks: the exchange bit for bit, 2 RK4 steps to 1e-12 relative L2."""
    import numpy as np
    from oracle import port, ref
    n, rank = pool.size(), pool.rank()
    blk = (32, 8, 8)
    sid = 1 if scheme == "hybrid" else 0
    mu = 1e-2
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    flux = make_flux(sp, gas, scheme, mu)
    L = 2 * np.pi
    if cfgid == 5:
        fix = np.load(os.path.join(ROOT, "tests", "golden", "config5_amr.npz"))
        boxes = fix["p_boxes"]
        nb = tuple(int(x) for x in fix["p_roots"])
        nlb_glob = len(boxes)
        first, cnt = contiguous_partition(nlb_glob, n, rank)
        grid = sp.cartesian_grid_t.from_boxes(blk, boxes[first:first + cnt], pool, first_block=first)
        tables = amr_tables(fix, "p", n, rank, np)
        q0 = np.zeros((nlb_glob, blk[2] + 2 * NG, blk[1] + 2 * NG, blk[0] + 2 * NG, 5))
        for lb in range(nlb_glob):
            ax = [boxes[lb, 2 * d] + (np.arange(-NG, blk[d] + NG) + 0.5) * (boxes[lb, 2 * d + 1] - boxes[lb, 2 * d]) / blk[d] for d in range(3)]
            Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
            q0[lb, ..., 0] = P0 * (1 + 0.05 * np.cos(X) * np.cos(Y))
            q0[lb, ..., 1] = T0 * (1 + 0.02 * np.sin(X + 2 * Y - Z))
            q0[lb, ..., 2] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
            q0[lb, ..., 3] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
            q0[lb, ..., 4] = 10.0 * np.sin(Z) * np.cos(X + Y)
        q0 *= 1 + 1e-2 * np.random.default_rng(61).uniform(-1, 1, q0.shape)
        send_all = np.concatenate([fix[f"p_send_{n}_{r}"] for r in range(n)], axis=0).astype(np.int64)
        isend_all = np.concatenate([fix[f"p_isend_{n}_{r}"] for r in range(n)], axis=0).astype(np.int64)
        port.set_amr(boxes, send_all, isend_all)
        desc = f"AMR 2x2x2 roots of {blk[0]}x{blk[1]}x{blk[2]} cells, root 0 refined: {nlb_glob} blocks, tables of the reference for {n} ranks"
        dxmin = float((boxes[:, 1] - boxes[:, 0]).min()) / blk[0]
    else:
        nb = (2, 2, 2 * n)
        blocks = sp.cartesian_blocks_t(nb, [0.0, L] * 3)
        grid = sp.cartesian_grid_t(blk, blocks, sp.identity(), pool)
        first, cnt = grid.first_block, grid.num_local_blocks
        tables = None
        q0 = host_state(nb, np, block=blk, perturb=1e-2, seed=31)
        desc = f"{nb[0]}x{nb[1]}x{nb[2]} blocks of {blk[0]}x{blk[1]}x{blk[2]} cells, periodic, perturbed TGV"
        dxmin = min(L / (nb[d] * blk[d]) for d in range(3))
    try:
        cfg = ref.make_cfg(nb, blk, NG, periodic=(1, 1, 1), scheme=sid, gamma=GAMMA, R=RGAS, mu=mu, prandtl=PRANDTL, sensor_eps=1e-2,
                           nranks=1, integrator=0)
        # (a) exchange alone, from zeroed ghosts: bit for bit
        qz = q0.copy()
        mask = np.zeros(q0.shape[1:4], dtype=bool)
        mask[NG:-NG, NG:-NG, NG:-NG] = True
        qz[:, ~mask, :] = 0.0
        want_ex = port.exchange(cfg, qz.ravel()).reshape(q0.shape)
        qa = sp.grid_array.from_host(grid, qz[first:first + cnt])
        ex = sp.make_exchange(qa, (True, True, True), tables=tables)
        ex.exchange(qa)
        exch_ok = bool(np.array_equal(qa.to_host(), want_ex[first:first + cnt]))
        # (b) 2 RK4 steps through the fused, overlapped path
        qe = port.exchange(cfg, q0.ravel()).reshape(q0.shape)
        dt = 0.2 * dxmin / port.reduce_umax(cfg, qe.ravel())
        want = port.advance(cfg, qe.ravel(), dt, 2).reshape(q0.shape)
        qa = sp.grid_array.from_host(grid, qe[first:first + cnt])
        ex2 = sp.make_exchange(qa, (True, True, True), tables=tables)
        ti = sp.integrator_t(sp.time_axis_t(0.0, dt), sp.rk4_t, sp.integrator_data_t(qa, sp.grid_array(grid, 0.0), sp.rk4_t),
                             sp.flux_div_rhs_t(flux, sp.overwrite), sp.exchange_bc_t(ex2), sp.state_transform_t(gas))
        for _ in range(2):
            ti.advance()
        got = ti.solution().to_host()
        ref_slab = want[first:first + cnt]
        num, den = float(((got - ref_slab) ** 2).sum()), float((ref_slab ** 2).sum())
    finally:
        if cfgid == 5:
            port.set_amr()
    red = pool.reduce(num, sp.RED_SUM), pool.reduce(den, sp.RED_SUM)
    err = (red[0] / red[1]) ** 0.5
    exch_all = pool.reduce(0.0 if exch_ok else 1.0, sp.RED_SUM) == 0.0
    return {"ok": bool(exch_all and err < 1e-12), "exchange_bit_exact": bool(exch_all), "trajectory_rel_l2": err, "tolerance": 1e-12,
            "steps": 2, "grid": desc, "ranks": n, "functor_set": "hybrid(totani_lr,fweno_t,ducros_t) + visc_lr" if sid else "totani_lr + visc_lr",
            "path": {"fused_stage_kernel": ti._plan is not None, "ghosts_in_kernel": bool(ti._fuse_exchange),
                     "messages": ("peer memory (CUDA IPC)" if ex2._p2p else "NCCL send/recv") if n > 1 else "none (1 rank)",
                     "two_streams": bool(ti._two_streams and n > 1)},
            "oracle": "oracle.port (C restatement, pinned against the unmodified reference in tests/test_oracle_vs_reference.py)"}


# ------------------------------------------------------------------------------------------------------
def e2e_leg(w, args, sp, torch, dist, world, steps):
    """End to end through the public API with HOST buffers: every step the state comes from pinned host memory (H2D), the step
    runs, and the new state goes back to pinned host memory (D2H) together with the max-wavespeed scalar the CFL logic reads
    (development/cuda-tgv/main.cc:228). Three streams: the upload of step s+1 and the download of step s-1 run while step s
    computes; every byte still crosses PCIe inside the timed region."""
    q = w.q
    host_in = torch.empty(q.data.shape, dtype=torch.float64, pin_memory=True)
    host_out = torch.empty(q.data.shape, dtype=torch.float64, pin_memory=True)
    host_in.copy_(q.data)
    torch.cuda.synchronize()
    ksteps = max(2, min(steps, 5))
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    bufs = [q.data, torch.empty_like(q.data)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    drained = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    up.wait_stream(main)
    with torch.cuda.stream(up):
        bufs[0].copy_(host_in, non_blocking=True)
        ready[0].record()
    for s_ in range(ksteps):
        cur = s_ % 2
        if s_ + 1 < ksteps:
            with torch.cuda.stream(up):
                if s_ >= 1:
                    up.wait_event(drained[1 - cur])          # the result of step s-1 has left bufs[1 - cur]
                bufs[1 - cur].copy_(host_in, non_blocking=True)
                ready[1 - cur].record()
        main.wait_event(ready[cur])
        q.data = bufs[cur]
        w.ti.advance()
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(down):
            down.wait_event(done)
            host_out.copy_(q.data, non_blocking=True)
            drained[cur].record()
        sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, w.gas)     # D2H of the scalar + cross-rank max (synchronises the step)
    main.wait_stream(down)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ems = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ems], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ems = float(tt.item())
    nbytes = int(host_in.numel() * 8)
    q.data = bufs[0]
    out = {"value": w.cells_total * STAGES * ksteps / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nbytes,
           "d2h_bytes_per_step": nbytes + 8, "steps": ksteps, "ms_per_step": ems / ksteps,
           "pcie_GBps_each_way": nbytes / (ems / ksteps * 1e-3) / 1e9,
           "note": "per step: state H2D from pinned host memory, one RK4 step through integrator_t.advance(), new state D2H into pinned host "
                   "memory + the max-wavespeed scalar; uploads, compute and downloads of neighbouring steps overlap on three streams"}
    del host_in, host_out, bufs
    return out


def ref_gpu_baseline():
    """The reference's own CUDA path on this GPU: integration/_build/ref_gpu_bench is the unmodified reference
    compiled with nvcc -arch=sm_100a in the dev container (device::gpu arrays, tags `basic` / `fldbc` / `fused`). Rank 0, N = 1."""
    exe = os.path.join(ROOT, "integration", "_build", "ref_gpu_bench")
    if not os.path.exists(exe):
        return {"unavailable": "integration/_build/ref_gpu_bench not built (needs /root/reference at build time)"}
    out = {}
    for scheme, name in ((0, "central"), (1, "hybrid")):
        try:
            r = subprocess.run([exe, "8", "32", "2", str(scheme)], capture_output=True, text=True, timeout=240)
            rec = json.loads(r.stdout.strip().splitlines()[-1])
            out[name] = rec
        except Exception as exc:
            out[name] = {"error": str(exc)[:200]}
    return out


def ours(args):
    import torch
    import torch.distributed as dist
    import spade_b200.api as sp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; spade_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL's send/recv kernels on a high-priority stream: they need a few CTA slots while the stage kernel fills every SM,
        # and at normal priority they would only be scheduled once that grid drains (no overlap)
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = os.environ.get("SPB_NCCL_HIGH_PRIO", "1") != "0"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    pool = sp.pool_t.from_torch()
    if os.environ.get("SPB_DIAG_SOLO"):
        pool = sp.pool_t()                   # diagnosis: every rank runs the 1-GPU workload on its own (process group initialised, no exchange)
    n = max(world, 1)
    timing = {"on": False, "events": []}

    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(args.config, args.scheme, sp, pool, torch)
        except Exception as exc:             # a broken checker must be visible, not fatal to the measurement
            parity = {"ok": False, "error": repr(exc)[:300]}

    w = build_workload(args.config, args, sp, pool, torch, timing)
    m = measure(w, args.steps, args.warmup, sp, torch, dist, world, rank, local_rank, timing, True)
    roofline = roofline_of(w, m, args.steps)

    # the RHS alone (pde_algs::flux_div with the overwrite trait, 80 B per cell), timed after the run for the record
    for _ in range(2):
        sp.flux_div(w.q, w.rhs, w.flux, sp.overwrite)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    r0.record()
    for _ in range(5):
        sp.flux_div(w.q, w.rhs, w.flux, sp.overwrite)
    r1.record()
    torch.cuda.synchronize()
    rhs_ms = r0.elapsed_time(r1) / 5
    peak, _ = measured_peak_hbm()
    roofline["rhs_only"] = {"kernel": "flux_div (overwrite)", "alg_bytes_per_cell": 80.0, "ms_per_launch": rhs_ms,
                            "achieved": 80.0 * w.cells_local / (rhs_ms * 1e-3) / 1e9,
                            "frac": 80.0 * w.cells_local / (rhs_ms * 1e-3) / 1e9 / peak}
    if w.scheme == "hybrid":
        tf = HYBRID_FLOP_PER_CELL * w.cells_local / (rhs_ms * 1e-3) / 1e12
        roofline["rhs_only"].update({"fp64_tflops": tf, "fp64_frac": tf / FP64_PEAK_NOMINAL_TFLOPS})
    tpc, tsrc = measured_traffic("rhs")
    if tpc and w.scheme == "central":
        roofline["rhs_only"]["traffic"] = tpc * w.cells_local

    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(w, args, sp, torch, dist, world, args.steps)

    # The other BASELINE configurations in the same line of the default run, so that the driver's records carry them at every N:
    # config 4 (256^3 per GPU, the weak-scaling sweep BASELINE names), config 3 (stretched channel, hybrid WENO + Ducros + viscous,
    # a 1024 x 512 x 64 slab per GPU: N = 8 is the 1024 x 512 x 512 grid) and config 5 (the AMR grid, strong scaling), each with
    # its own parity check through the same code path. A sub-record that fails reports the error; the main line stands.
    configs = None
    if args.config == 2 and not args.no_configs and not args.lattice:
        configs = {}
        for cid in (4, 3, 5):
            try:
                sub_scheme = "hybrid" if cid == 3 else "central"
                sub_args = argparse.Namespace(**{**vars(args), "config": cid, "scheme": sub_scheme})
                ksteps = max(args.steps, 20) if cid == 4 else max(5, min(args.steps, 10))
                par = None
                if cid != 4 and not args.no_parity:          # config 4 runs the functor set and code path the main line has checked
                    try:
                        par = parity_check(cid, sub_scheme, sp, pool, torch)
                    except Exception as exc:
                        par = {"ok": False, "error": repr(exc)[:300]}
                wc = build_workload(cid, sub_args, sp, pool, torch, timing)
                mc = measure(wc, ksteps, max(args.warmup, 3), sp, torch, dist, world, rank, local_rank, timing, False)
                rc = roofline_of(wc, mc, ksteps)
                keys = ("bound", "achieved", "peak", "unit", "frac", "ms_per_launch", "alg_bytes_per_cell", "step_share", "frac_of_measured_dfma_rate", "hbm")
                configs[f"config{cid}"] = {"workload": workload_config(sub_args, n)["workload"],
                                           "value": mc["value"], "unit": UNIT, "ms_per_step": mc["ms"] / ksteps, "steps": ksteps,
                                           "gpu_launches": mc["launches"], "scaling": wc.scaling,
                                           "roofline": {k: rc[k] for k in keys if k in rc}}
                if par is not None:
                    configs[f"config{cid}"]["parity_check"] = par
                del wc
            except Exception as exc:                         # never take the headline down with a sub-record
                configs[f"config{cid}"] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()

    cpu_baseline, ref_gpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_sample(args, 2)
            cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "flags")}
        except Exception as exc:     # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "error": str(exc)}
        ref_gpu = ref_gpu_baseline()
        if isinstance(ref_gpu, dict) and "unavailable" not in ref_gpu:
            # our kernels on the same grid as the reference-GPU run (256^3 in 32^3 blocks), for the ratio
            pass

    if rank == 0:
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT,
                "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps,
                "higher_is_better": True, "scaling": w.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, n), "cell_steps_per_s": m["value"] / STAGES,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": m["launches"],
                "clocks": m["clocks"], "parity_check": parity}
        line["steps_detail"] = m["steps_detail"]
        if m.get("phases"):
            line["phases"] = m["phases"]
        if configs:
            line["configs"] = configs
        if ref_gpu is not None:
            line["ref_gpu_baseline"] = ref_gpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
