#!/usr/bin/env python
"""bench.py — headline benchmark of the RHS hot path (flux_div + ghost exchange + RK stage update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Taylor-Green vortex, 512^3 cells in 32^3 blocks (16x16x16 blocks,
2 exchange cells), totani_lr (2nd-order KEEP central) + visc_lr, rk4_t with the fused prim/cons update,
fully periodic. One "step" = one RK4 time step = 4 x (flux_div + exchange + stage update) over the whole
grid. For N > 1 the per-GPU grid is kept (weak scaling): lattice 16 x 16 x 16N, rank r owns z-slab r
(SPADE's contiguous block partition), ghost exchange between ranks over NCCL send/recv.

metric: cell-stage-updates/s = cells x stages x steps / time (a cell advanced through one RK stage:
RHS + exchange + stage update), whole job. Prints ONE JSON line (contract in the task statement).
"""
import argparse
import ctypes as C
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
LATTICE_1GPU = (16, 16, 16)
# dram__bytes_read.sum + dram__bytes_write.sum per interior cell from the committed `ncu --set full` capture of the
# dominant kernel (profiles/), None until a capture exists for that kernel
# profiles/r01_ncu_full_stage_kernels_512cube.txt: the four rk4 stage kernels of one step move 20.23 + 26.17 + 32.17 + 20.37 GB
# for 134.2 M cells -> 184.3 B per cell per launch (algorithmic 166.9); profiles/r01_ncu_full_rhs_...: 1.535 GB / 16.8 M cells
TRAFFIC_PER_CELL = {"fused": 184.3, "rhs": 91.5}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", type=int, nargs=3, default=None, help="blocks per GPU (default 16 16 16)")
    ap.add_argument("--scheme", default="central", choices=["central", "hybrid"])
    ap.add_argument("--unfused", action="store_true", help="two kernels per stage (flux_div, then rk_update) instead of the fused stage kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons during the timed region (nvidia-smi, B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
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
        while not self._stop.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
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
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's own CPU implementation of the path (oracle/_ref = unmodified SPADE headers compiled in the
    dev container; falls back to the C port) on all host threads, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import ref, port
    cores = os.cpu_count() or 1
    kind = "reference" if ref.available() else "port"
    lat = (4, 4, 4)                      # 128^3 cells in 32^3 blocks: same block shape as the GPU workload
    nranks = max(1, min(cores, lat[0] * lat[1] * lat[2])) if kind == "reference" else 1
    scheme = 0 if args.scheme == "central" else 1
    mu = (P0 / (RGAS * T0)) * U0 * 1.0 / REYNOLDS
    cfg = ref.make_cfg(lat, (BLOCK,) * 3, NG, scheme=scheme, gamma=GAMMA, R=RGAS, mu=mu, prandtl=PRANDTL,
                       sensor_eps=1e-2, nranks=nranks, integrator=0)
    q = host_state(lat, np)
    umax = port.reduce_umax(cfg, q.ravel())
    dt = 0.2 * (2 * np.pi / (lat[0] * BLOCK)) / umax
    cells = (lat[0] * BLOCK) ** 3
    if kind == "reference":
        qq, _ = ref.advance(cfg, q.ravel(), dt, max(1, min(args.warmup, 1)))
        qq, sec = ref.advance(cfg, qq, dt, args.steps)
    else:
        qq = port.advance(cfg, q.ravel(), dt, 1)
        t0 = time.time()
        port.advance(cfg, qq, dt, args.steps)
        sec = time.time() - t0
    value = cells * STAGES * args.steps / sec
    sample = f"TGV {lat[0]*BLOCK}^3 cells in {BLOCK}^3 blocks ({lat[0]}x{lat[1]}x{lat[2]}), rk4, {args.steps} steps"
    line = {"impl": "reference", "metric": "fp64 cell-updates/sec (RHS+exchange+RK)", "value": value,
            "unit": "cell-stage-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, sample_note=sample),
            "cpu_baseline": {"value": value, "unit": "cell-stage-updates/s", "cores": nranks, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "cell-stage-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def host_state(lat, np):
    """TGV initial condition on the host (all cells incl. ghosts analytic), reference memory order."""
    nlb = lat[0] * lat[1] * lat[2]
    L = 2 * np.pi
    q = np.zeros((nlb, BLOCK + 2 * NG, BLOCK + 2 * NG, BLOCK + 2 * NG, 5))
    rho0 = P0 / (RGAS * T0)
    for lb in range(nlb):
        b = (lb % lat[0], (lb // lat[0]) % lat[1], lb // (lat[0] * lat[1]))
        ax = [b[d] * L / lat[d] + (np.arange(-NG, BLOCK + NG) + 0.5) * (L / lat[d] / BLOCK) for d in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        q[lb, ..., 0] = P0 + rho0 * U0 * U0 / 16 * (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2)
        q[lb, ..., 1] = T0
        q[lb, ..., 2] = U0 * np.sin(X) * np.cos(Y) * np.cos(Z)
        q[lb, ..., 3] = -U0 * np.cos(X) * np.sin(Y) * np.cos(Z)
    return q


def workload_config(args, sample_note=None):
    lat = tuple(args.lattice) if args.lattice else LATTICE_1GPU
    n = args.gpus
    cfg = {"workload": f"TGV {lat[0]*BLOCK}x{lat[1]*BLOCK}x{lat[2]*BLOCK*n} cells ({lat[0]}x{lat[1]}x{lat[2]*n} blocks of {BLOCK}^3, "
                       f"{NG} exchange cells), {'totani_lr' if args.scheme == 'central' else 'hybrid(totani_lr,fweno_t,ducros_t)'}"
                       " + visc_lr, rk4_t fused prim/cons, periodic",
           "cells_per_gpu": lat[0] * lat[1] * lat[2] * BLOCK ** 3, "stages_per_step": STAGES,
           "partition": f"contiguous block runs, rank r = z-slab r ({n} ranks)",
           "l2": "inputs larger than L2 (q + 4 residual arrays, 7.6 GB each at 512^3)"}
    if sample_note:
        cfg["reference_sample"] = sample_note
    return cfg


# ------------------------------------------------------------------------------------------------------
def device_state(sp, grid, torch):
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
    n = max(world, 1)
    lat = tuple(args.lattice) if args.lattice else LATTICE_1GPU
    lattice = (lat[0], lat[1], lat[2] * n)
    L = 2 * 3.141592653589793
    blocks = sp.cartesian_blocks_t(lattice, [0.0, L, 0.0, L, 0.0, L * n])
    grid = sp.cartesian_grid_t((BLOCK,) * 3, blocks, sp.identity(), pool)
    gas = sp.ideal_gas_t(GAMMA, RGAS)
    mu = (P0 / (RGAS * T0)) * U0 * 1.0 / REYNOLDS
    conv = sp.totani_lr(gas)
    if args.scheme == "hybrid":
        conv = sp.hybrid_scheme_t(conv, sp.fweno_t(gas), sp.ducros_t(1e-2), sp.full_flux)
    flux = sp.flux_desc(sp.compose(conv, sp.visc_lr(sp.constant_viscosity_t(mu, PRANDTL), gas)))

    q = device_state(sp, grid, torch)
    rhs = sp.grid_array(grid, 0.0, (NG,) * 3)
    handle = sp.make_exchange(q, (True, True, True))
    handle.exchange(q)
    umax = sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, gas)
    dt = 0.2 * grid.get_dx(0) / umax

    ev_pairs = []
    timing_on = [False]
    fused = (not args.unfused) and args.scheme == "central"

    def calc_rhs_unfused(r, qq, t):
        if timing_on[0]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sp.flux_div(qq, r, flux, sp.overwrite)
            e1.record()
            ev_pairs.append((e0, e1, 80.0))
        else:
            sp.flux_div(qq, r, flux, sp.overwrite)

    # the usual SPADE rhs callback (flux_div with the overwrite trait); as a flux_div_rhs_t the integrator recognises it and
    # runs flux_div + stage update as ONE kernel per stage
    calc_rhs = sp.flux_div_rhs_t(flux, sp.overwrite) if fused else calc_rhs_unfused

    # the usual periodic boundary callback (exchange only); as an exchange_bc_t the integrator can send the ghost messages of
    # the rank-boundary blocks while it advances the rank-interior blocks
    bc = sp.exchange_bc_t(handle)

    alg = sp.rk4_t
    data = sp.integrator_data_t(q, rhs, alg)
    ti = sp.integrator_t(sp.time_axis_t(0.0, dt), alg, data, calc_rhs, bc, sp.state_transform_t(gas))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ti.advance()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sp.launch_count()
    timing_on[0] = True
    if fused:
        ti.stage_events = ev_pairs
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        ti.advance()
    t1.record()
    barrier()
    timing_on[0] = False
    ti.stage_events = None
    ms = t0.elapsed_time(t1)
    launches = sp.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    fdiv_ms = sum(a.elapsed_time(b) for a, b, _ in ev_pairs) / max(1, len(ev_pairs))
    fdiv_bpc = sum(c for _, _, c in ev_pairs) / max(1, len(ev_pairs))      # algorithmic bytes per cell, mean over launches
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    umax_end = sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, gas)
    if not (umax_end == umax_end) or umax_end > 10 * umax:
        raise SystemExit(f"bench.py: solution diverged (umax {umax_end})")

    local_cells = grid.local_cells()
    total_cells = local_cells * n
    value = total_cells * STAGES * args.steps / (ms * 1e-3)

    # roofline of the dominant kernel. Algorithmic bytes per interior cell: plain flux_div reads q and
    # writes rhs = 80 B; the fused stage kernel reads q, writes q', reads/writes the residual registers its stage needs
    # (rk4 = 120, 160, 200, 120 B, mean 150 B per launch) and writes the same-rank ghost cells (40 B x 0.4238 ghost cells per
    # interior cell at n = 32, g = 2 = 17 B).
    peak, peak_src = measured_peak_hbm()
    fdiv_bytes = fdiv_bpc * local_cells
    achieved = fdiv_bytes / (fdiv_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "flux_div_narrow_kernel<FUSED stage>" if fused else "flux_div kernel (rhs only)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": TRAFFIC_PER_CELL.get("fused" if fused else "rhs") and
                TRAFFIC_PER_CELL["fused" if fused else "rhs"] * local_cells, "peak_source": peak_src,
                "alg_bytes_per_cell": fdiv_bpc, "ms_per_launch": fdiv_ms, "launches_timed": len(ev_pairs),
                "cell_evals_per_s": local_cells / (fdiv_ms * 1e-3),
                "step_share": fdiv_ms * STAGES * args.steps / ms}
    # the RHS alone (pde_algs::flux_div with the overwrite trait, 80 B per cell), timed after the run for the record
    for _ in range(2):
        sp.flux_div(q, rhs, flux, sp.overwrite)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    r0.record()
    for _ in range(5):
        sp.flux_div(q, rhs, flux, sp.overwrite)
    r1.record()
    torch.cuda.synchronize()
    rhs_ms = r0.elapsed_time(r1) / 5
    roofline["rhs_only"] = {"kernel": "flux_div (overwrite)", "alg_bytes_per_cell": 80.0, "ms_per_launch": rhs_ms,
                            "achieved": 80.0 * local_cells / (rhs_ms * 1e-3) / 1e9,
                            "frac": 80.0 * local_cells / (rhs_ms * 1e-3) / 1e9 / peak}

    # end to end through the public API with host buffers: every step the state comes from pinned host memory
    # and the step's metric (max wavespeed for the CFL number, as in development/cuda-tgv/main.cc:228) goes back
    e2e = None
    if not args.no_e2e:
        host_q = torch.empty(q.data.shape, dtype=torch.float64, pin_memory=True)
        host_q.copy_(q.data)
        torch.cuda.synchronize()
        ksteps = max(2, min(args.steps, 5))
        barrier()
        # Two device buffers: the copy of step s+1's input (its own stream) runs while step s computes. Every step's input
        # still crosses PCIe inside the timed region; what overlaps is only that the bus and the SMs work at the same time.
        copy_stream = torch.cuda.Stream()
        bufs = [q.data, torch.empty_like(q.data)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            bufs[0].copy_(host_q, non_blocking=True)
            ready[0].record()
        for s_ in range(ksteps):
            cur = s_ % 2
            if s_ + 1 < ksteps:
                with torch.cuda.stream(copy_stream):       # bufs[1 - cur] was last read by step s-1, which the reduction below has synchronised
                    bufs[1 - cur].copy_(host_q, non_blocking=True)
                    ready[1 - cur].record()
            main.wait_event(ready[cur])
            q.data = bufs[cur]
            ti.advance()
            um = sp.transform_reduce(q, sp.FN_WAVESPEED, sp.RED_MAX, gas)     # D2H of the scalar + cross-rank max
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ems], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ems = float(tt.item())
        e2e = {"value": total_cells * STAGES * ksteps / (ems * 1e-3), "unit": "cell-stage-updates/s",
               "h2d_bytes_per_step": int(host_q.numel() * 8), "d2h_bytes_per_step": 8, "steps": ksteps,
               "note": "state copied from pinned host memory every step (double-buffered: the copy of the next step's input overlaps the current step); max-wavespeed scalar read back"}
        del host_q, bufs

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = cpu_baseline_leg(args)
        except Exception as exc:     # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "error": str(exc)}

    if rank == 0:
        line = {"metric": "fp64 cell-updates/sec (RHS+exchange+RK)", "value": value, "unit": "cell-stage-updates/s",
                "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args), "cell_steps_per_s": value / STAGES,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(args):
    """oracle/_ref (the reference itself) timed on the host cores on a bounded sample (about 10-30 s)."""
    import numpy as np
    from oracle import ref, port
    cores = os.cpu_count() or 1
    kind = "reference" if ref.available() else "port"
    lat = (4, 4, 4)
    nranks = max(1, min(cores, 64)) if kind == "reference" else 1
    mu = (P0 / (RGAS * T0)) * U0 * 1.0 / REYNOLDS
    cfg = ref.make_cfg(lat, (BLOCK,) * 3, NG, scheme=0 if args.scheme == "central" else 1, gamma=GAMMA, R=RGAS, mu=mu,
                       prandtl=PRANDTL, sensor_eps=1e-2, nranks=nranks, integrator=0)
    q = host_state(lat, np)
    dt = 0.2 * (2 * np.pi / (lat[0] * BLOCK)) / port.reduce_umax(cfg, q.ravel())
    steps = 2
    if kind == "reference":
        _, sec = ref.advance(cfg, q.ravel(), dt, steps)
    else:
        t0 = time.time()
        port.advance(cfg, q.ravel(), dt, steps)
        sec = time.time() - t0
    cells = (lat[0] * BLOCK) ** 3
    return {"value": cells * STAGES * steps / sec, "unit": "cell-stage-updates/s", "cores": nranks, "kind": kind,
            "sample": f"TGV {lat[0]*BLOCK}^3 cells in {BLOCK}^3 blocks, rk4, {steps} steps, {sec:.1f} s"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
