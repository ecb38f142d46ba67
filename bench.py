#!/usr/bin/env python
"""bench.py — converged PNJL gap points/s on N B200s (BASELINE.json metric), with roofline, CPU baseline, e2e.

    python bench.py --gpus N --steps K --warmup W [--workload cfg5|cfg4|cfg3|cfg2] [--impl reference]

A "step" is one full pass of the hot path over the workload grid:
  cfg5 (default; BASELINE.json configs[4], the config the north-star target is quoted on): the anisotropic
       1024(T) x 1024(mu) x 8(xi) scan, T 50..300 MeV, mu_q 0..400 MeV, xi in {-0.6..0.6 step 0.2, 0.8},
       64x16 Gauss-Legendre nodes, max_iter 40; per (mu, xi) line MultiSeed at T[0] then
       PhaseAwareContinuitySeed along T (run_gap_transport_scan.jl order).  8.39 M points, 8192 lines.
  cfg4: 2048x2048 window near the CEP (T 100..160, mu_q 260..330, xi=0), 64x16.
  cfg3: 256x256x8 independent points, MultiSeed at every point, 64x16.
  cfg2: 128x128 isotropic scan, 12x6 nodes (the script's defaults).
With N GPUs the mu axis is dealt out over the ranks (round-robin by default, --layout slab for contiguous slabs); each
rank runs its lines with no exchange and its kernel stores the records straight into rank 0's array over NVLink (or rank 0
gathers them over NCCL, --gather nccl), inside the timed region; `value` = all converged points / max-over-ranks device time.
  --scaling strong (default) the SAME grid at every N — BASELINE's "1024x1024x8 at 1/2/4/8 B200": n_mu / N mu-values per GPU.
                   At N > 1 a short weak-scaling measurement follows and is reported under the key "weak".
  --scaling weak   every GPU carries one full workload slab, i.e. the mu axis is refined to N x n_mu points over the same
                   range; N = 1 is exactly the BASELINE config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

XI8 = [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8]
WORKLOADS = {
    # name: (kind, xi list, n_mu, mu range, n_T, T range, p_num, t_num)
    "cfg5": ("lines", XI8, 1024, (0.0, 400.0), 1024, (50.0, 300.0), 64, 16),
    "cfg4": ("lines", [0.0], 2048, (260.0, 330.0), 2048, (100.0, 160.0), 64, 16),
    "cfg3": ("points", XI8, 256, (0.0, 400.0), 256, (50.0, 300.0), 64, 16),
    "cfg2": ("lines", [0.0], 128, (0.0, 400.0), 128, (50.0, 300.0), 12, 6),
}
DESCR = {
    "cfg5": "BASELINE configs[4]: anisotropic 1024x1024x8 (T,mu,xi) continuity scan, GL 64x16, max_iter 40",
    "cfg4": "BASELINE configs[3]: 2048x2048 scan near the CEP, xi=0, GL 64x16, max_iter 40",
    "cfg3": "BASELINE configs[2]: 256x256x8 independent points, MultiSeed everywhere, GL 64x16, max_iter 40",
    "cfg2": "BASELINE configs[1]: isotropic 128x128 T-mu continuity scan, GL 12x6, max_iter 40",
}
MAX_ITER = 40
HBARC = 197.327


FUSED_FLOP = 90.0   # fused final pass: residual F without second derivatives (≈ 61) + the thermo sums that do not
                    # share work with it (≈ 29); DESIGN.md §5


def alg_flops(n_nodes, n_fj, n_th, n_flav=3, n_ft=0.0):
    """123 FLOP per node x flavour of an Omega-gradient/Jacobian pass, 54 per thermo-pass unit.
    n_flav = flavours actually evaluated per node: 2 when the kernel uses M_u == M_d (isospin_symmetric, the
    default: every state on these grids has phi_u == phi_d), 3 for the reference's loop."""
    return n_nodes * n_flav * (123.0 * n_fj + 54.0 * n_th + FUSED_FLOP * n_ft)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def build_lines(w):
    kind, xis, n_mu, (m0, m1), n_T, (t0, t1), p, t = WORKLOADS[w]
    mus = np.linspace(m0, m1, n_mu)
    T = np.linspace(t0, t1, n_T)
    return xis, mus, T, p, t


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference algorithm, nested-dual AD like the reference) on all
# host cores, on a bounded sample of the same workload.
# ------------------------------------------------------------------------------------------------------------
def _lines_sample(workload, n_lines, offset=0):
    """`n_lines` complete lines of the workload grid spread over (xi, mu) by a golden-ratio sequence (every xi, mu over
    the whole range — what one rank of an interleaved multi-GPU run carries)."""
    xis, mus, T, p, t = build_lines(workload)
    n_mu = len(mus)
    total = len(xis) * n_mu
    li = ((((np.arange(n_lines) + offset) * 0.6180339887498949) % 1.0) * total).astype(int) % total
    return li, np.array([xis[i // n_mu] for i in li], dtype=float), mus[li % n_mu], T


def cpu_sample(workload, target_seconds=12.0, threads=None, engine="ad", step=0):
    """The path on the host CPU, SAME CONFIG as the GPU arm: lines of the workload grid marched over the true T grid with the
    reference's seeding (MultiSeed at T[0], then PhaseAwareContinuitySeed — run_gap_transport_scan.jl:417-443), `threads`
    lines in parallel.
      engine "analytic"  oracle/pnjl_analytic_cpu.cpp: the GPU kernel's own algorithm (closed-form Jacobian, isospin shortcut,
                         fused final pass) compiled for the host — the analytic-Jacobian CPU baseline of SURVEY.md §8d.  Fast
                         enough to march complete lines.
      engine "ad"        oracle/pnjl_oracle.cpp: the reference's arithmetic (Omega once, F and J by nested dual numbers like
                         ForwardDiff inside NLsolve's autodiff=:forward).  A complete 1024-point line takes ~35 s per core, so a
                         step marches one WINDOW of the T grid (a third of it; window `step` mod 3, so successive steps cover
                         the whole range) started from the TRUE previous solution of each line (computed by the analytic engine,
                         not timed): the seeds, and therefore the evaluations per point, are those of the complete line.  The
                         first window starts at T[0] with the MultiSeed bootstrap like a line does."""
    from oracle.oracle import Oracle, load_phase_tables
    from tests.hostsim.hostsim import HostSim
    kind, xis, n_mu, _, n_T, _, p, t = WORKLOADS[workload]
    xis, mus, T, p, t = build_lines(workload)
    threads = threads or len(os.sched_getaffinity(0))
    o = Oracle(p_num=p, t_num=t, max_iter=MAX_ITER, n_threads=threads)
    gdir = os.path.join(ROOT, "julia_relaxtime_b200", "data")
    tables, index = load_phase_tables(os.path.join(gdir, "boundary.csv"), os.path.join(gdir, "cep.csv"), xis)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    hs = HostSim(o.p_nodes, o.p_w, o.c_nodes, o.c_w, max_iter=MAX_ITER)
    rng = np.random.default_rng(step)
    conv_of = lambda rec: int(((rec[:, 21].astype(np.int64) & 1) != 0).sum())
    evals_of = lambda rec: int(rec[:, 22].sum() + rec[:, 30].sum())          # Jacobian passes + fused final passes
    if kind == "points":
        def run(n):
            Tm = rng.choice(T, n); mm = rng.choice(mus, n); xx = rng.choice(xis, n)
            t0 = time.perf_counter()
            if engine == "ad":
                r = o.solve_points(Tm / HBARC, mm / HBARC, xx, "multi")
                conv, nfj = int(r.converged.sum()), int(r.n_fj.sum())
            else:
                rec = hs.solve_points(Tm / HBARC, mm / HBARC, xx)
                conv, nfj = conv_of(rec), evals_of(rec)
            return time.perf_counter() - t0, conv, nfj
        dt, _, _ = run(threads)
        n = int(max(threads, min(400000, threads * target_seconds / max(dt, 1e-3))))
        dt, conv, nfj = run(n)
        return dict(points=n, seconds=dt, converged=conv, n_fj=nfj, threads=threads, engine=engine,
                    sample="%d random grid points of the %s grid, MultiSeed each" % (n, workload))

    if engine == "analytic":
        def run(n_lines, n_t):
            li, lx, lm, _ = _lines_sample(workload, n_lines, offset=step * n_lines)
            tidx = np.array([index[x] for x in lx], dtype=np.int32)
            t0 = time.perf_counter()
            rec = hs.scan_lines(lm, lx, T[:n_t], tables, tidx).reshape(-1, 32)
            return time.perf_counter() - t0, rec.shape[0], conv_of(rec), evals_of(rec)
        dt, tot, _, _ = run(threads, min(n_T, 64))
        per_line = dt * n_T / min(n_T, 64) / threads
        n_lines = int(max(threads, min(4096, threads * round(target_seconds / max(per_line * threads, 1e-3)))))
        dt, tot, conv, nfj = run(n_lines, n_T)
        return dict(points=tot, seconds=dt, converged=conv, n_fj=nfj, threads=threads, engine=engine,
                    sample="%d complete lines (spread over xi and mu) x all %d T of the %s grid" % (n_lines, n_T, workload))

    n_win = 3 if n_T >= 96 else 1
    wlen = (n_T + n_win - 1) // n_win
    w0 = (step % n_win) * wlen
    w1 = min(n_T, w0 + wlen)

    def run(groups, t_lo, t_hi):
        n_lines = groups * threads
        li, lx, lm, _ = _lines_sample(workload, n_lines, offset=step * n_lines)
        tidx = np.array([index[x] for x in lx], dtype=np.int32)
        init = None
        if t_lo > 0:     # the true state of every line at T[t_lo - 1] (analytic engine, untimed)
            pre = hs.scan_lines(lm, lx, T[:t_lo], tables, tidx)
            init = pre[:, -1, 0:5].copy()
        t0 = time.perf_counter()
        r = o.scan_lines(lm, lx, T[t_lo:t_hi], tables, tidx, init_x=init, init_T_MeV=float(T[t_lo - 1]) if t_lo > 0 else 0.0)
        return time.perf_counter() - t0, r.n, int(r.converged.sum()), int(r.n_fj.sum())

    dt, tot, _, _ = run(1, w0, min(w1, w0 + 24))
    per_group = dt * (w1 - w0) / max(1, min(w1, w0 + 24) - w0)
    groups = int(max(1, min(16, round(target_seconds / max(per_group, 1e-3)))))
    dt, tot, conv, nfj = run(groups, w0, w1)
    return dict(points=tot, seconds=dt, converged=conv, n_fj=nfj, threads=threads, engine=engine,
                sample="%d lines (spread over xi and mu) x the T window [%d, %d) of the %d-point T grid of %s (%.1f-%.1f MeV), "
                       "%d lines at a time, each line started from its true previous solution%s" % (
                           groups * threads, w0, w1, n_T, workload, T[w0], T[w1 - 1], threads,
                           " (window 0: MultiSeed bootstrap at T[0])" if w0 == 0 else ""))


def config_of(w, world, scaling, layout, extra=None):
    """The `config` object both arms print (same keys, so that the driver can tell they ran the same workload)."""
    kind, xis, n_mu, mr, n_T, tr, p, t = WORKLOADS[w]
    c = {"workload": w, "description": DESCR[w], "points": int(len(xis) * n_mu * n_T), "lines": int(len(xis) * n_mu) if kind == "lines" else 0,
         "n_T": int(n_T), "n_mu": int(n_mu), "xi": list(xis), "nodes": "%dx%d" % (p, t), "max_iter": MAX_ITER,
         "seeding": ("MultiSeed at T[0] of every (xi, mu) line, then PhaseAwareContinuitySeed along T" if kind == "lines"
                     else "MultiSeed at every point"),
         "nodes_note": "64- and 16-point Gauss-Legendre rules by Newton iteration on the Legendre recurrence (the reference's "
                       "FastGaussQuadrature switches to asymptotic formulas above n = 60; no full-precision golden output exists "
                       "at this mesh, SURVEY.md §8c) — results at this mesh are pinned to the oracle, not to Julia output"}
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    """--impl reference: the reference algorithm on the host CPU (oracle port; Julia is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = args.workload
    for k in range(args.warmup):
        cpu_sample(w, target_seconds=1.0, step=k)
    tot_pts = tot_s = tot_ev = tot_n = 0.0
    last = None
    for k in range(args.steps):
        last = cpu_sample(w, target_seconds=args.cpu_seconds, step=k)
        tot_pts += last["converged"]
        tot_s += last["seconds"]
        tot_ev += last["n_fj"]
        tot_n += last["points"]
    val = tot_pts / tot_s
    out = {"impl": "reference", "metric": "converged PNJL gap points/sec", "value": val, "unit": "points/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_of(w, 1, args.scaling, args.layout),
           "cpu_baseline": {"value": val, "unit": "points/s", "cores": last["threads"], "kind": "port",
                            "sample": last["sample"], "evals_per_point": tot_ev / max(1.0, tot_n),
                            "note": "C++ restatement of the reference algorithm (nested-dual AD Jacobian like "
                                    "ForwardDiff, NLsolve-faithful Newton/dogleg), OpenMP over lines; the Julia "
                                    "reference itself is single-threaded and cannot run in this image"},
           "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling measurement at N > 1")
    ap.add_argument("--gather", choices=("peer", "nccl"), default="peer",
                    help="multi-GPU result collection: peer = every rank's kernel stores its records straight into rank 0's buffer "
                         "over NVLink (CUDA IPC; falls back to nccl if the mapping cannot be set up); nccl = dist.gather after the kernel")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--layout", default="interleaved", choices=["interleaved", "slab"],
                    help="how the mu axis is dealt out over the ranks (results identical)")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps (ncu traffic captures only)")
    ap.add_argument("--n-mu", type=int, default=0, help="override the mu density (profiling runs only; recorded in config)")
    ap.add_argument("--n-t", type=int, default=0, help="override the T density (profiling runs only; recorded in config)")
    args = ap.parse_args()
    if args.n_mu or args.n_t:
        k, xs, nm, mr, nt, tr, pp, tt = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (k, xs, args.n_mu or nm, mr, args.n_t or nt, tr, pp, tt)
        DESCR[args.workload] += " [REDUCED GRID for profiling: n_mu=%d n_T=%d]" % (args.n_mu or nm, args.n_t or nt)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from julia_relaxtime_b200 import _abi as A
    from julia_relaxtime_b200._lib import Engine
    from julia_relaxtime_b200.distributed import PeerRecords, rank_line_indices, scan_sharded
    from julia_relaxtime_b200.scan import build_grid

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = args.workload
    kind = WORKLOADS[w][0]
    base_workload = WORKLOADS[w]
    p, t = base_workload[6], base_workload[7]
    n_nodes = p * t
    all_iso = all(x == 0.0 for x in base_workload[1])
    eng = Engine(p_num=p, t_num=t, max_iter=MAX_ITER, device=local_rank, lanes_per_solve=args.lanes)
    if all_iso and args.lanes == 0:
        eng.set_option("isotropic_batch", 1)     # the device entry points cannot look at xi: size the teams for p_num nodes
    stream = torch.cuda.current_stream().cuda_stream
    peak_burst, peak_sus = eng.measure_fp64_peak(1.0)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev)   # 512 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def checksum(rec2d):
        """Order-independent fingerprint of a record array: the 64-bit patterns added with wrap-around."""
        return rec2d.contiguous().view(torch.int64).sum()

    def run_job(mu_factor, steps, warmup, sample_clocks):
        """One measurement of the workload with the mu axis refined mu_factor-fold (1 = the BASELINE grid).  Returns a dict."""
        kind_, xs_, nm_, mr_, nt_, tr_, pp_, tt_ = base_workload
        WORKLOADS[w] = (kind_, xs_, nm_ * mu_factor, mr_, nt_, tr_, pp_, tt_)
        xis, mus, T, _, _ = build_lines(w)
        n_mu, n_T = len(mus), len(T)
        J = {}
        peer = None
        if kind == "lines":
            grid = build_grid(xis, 3.0 * mus, T)          # muB = 3 muq
            eng.set_boundaries(grid.tables)
            mine = rank_line_indices(len(xis), n_mu, rank, world, args.layout)
            d_muq = torch.as_tensor(grid.muq_MeV[mine], device=dev)
            d_xi = torch.as_tensor(grid.xi[mine], device=dev)
            d_tidx = torch.as_tensor(grid.table_idx[mine], device=dev)
            d_T = torch.as_tensor(grid.T_MeV, device=dev)
            d_out = torch.as_tensor(mine.astype(np.int64), device=dev)
            d_rec = torch.empty((len(mine), n_T, A.REC_DOUBLES), dtype=torch.float64, device=dev)
            n_total = grid.n_lines * n_T

            def compute(_lines=None):
                eng.scan_lines_device(d_muq, d_xi, d_tidx, d_T, d_rec, stream)
                return d_rec

            if world > 1 and args.gather == "peer":
                try:
                    peer = PeerRecords(grid.n_lines, n_T, rank, world, dev)
                except Exception as exc:                                     # noqa: BLE001 - any failure -> NCCL gather
                    if rank == 0:
                        print("peer gather unavailable (%s); using dist.gather" % exc, file=sys.stderr)
                    peer = None
            full_holder = [None]

            def timed_step(e1):
                if peer is not None:
                    # the kernel writes into rank 0's array itself; e1 = this rank's kernel, then only the closing barrier
                    eng.scan_lines_device_indexed(d_muq, d_xi, d_tidx, d_T, peer.ptr, d_out, stream)
                    e1.record()
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                else:
                    compute()
                    e1.record()
                    if world > 1:
                        full_holder[0], _ = scan_sharded(grid, len(xis), n_mu, lambda _l: d_rec, rank, world, layout=args.layout)
        else:
            gx, gm, gT = np.meshgrid(np.asarray(xis), mus, T, indexing="ij")
            allp = np.stack([gT.ravel() / HBARC, gm.ravel() / HBARC, gx.ravel()], axis=0)
            n_total = allp.shape[1]
            if n_total % world:
                raise SystemExit("points workload must divide evenly over ranks")
            idx = np.arange(rank, n_total, world)              # interleaved over the ranks like the lines
            d_T = torch.as_tensor(allp[0, idx].copy(), device=dev)
            d_mu = torch.as_tensor(allp[1, idx].copy(), device=dev)
            d_xi = torch.as_tensor(allp[2, idx].copy(), device=dev)
            d_rec = torch.empty((len(idx), A.REC_DOUBLES), dtype=torch.float64, device=dev)
            gather_bufs = [torch.empty_like(d_rec) for _ in range(world)] if (world > 1 and rank == 0) else None
            J["host_inputs"] = (allp, idx)

            def compute(_lines=None):
                eng.solve_points_device(d_T, d_mu, d_xi, d_rec, A.SEED_MULTI, None, 6, stream)
                return d_rec

            def timed_step(e1):
                compute()
                e1.record()
                if world > 1:
                    dist.gather(d_rec, gather_bufs if rank == 0 else None, dst=0)

        for _ in range(warmup):
            timed_step(torch.cuda.Event(enable_timing=True))
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            sampler.start()
        step_ms, kern_ms = [], []
        for _ in range(steps):
            if not args.no_flush:
                flush.fill_(1.0)                           # L2 flush between timed iterations (not timed)
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            timed_step(e1)
            e2.record()
            barrier()
            step_ms.append(e0.elapsed_time(e2))
            kern_ms.append(e0.elapsed_time(e1))
        J["clocks"] = sampler.stop() if (rank == 0 and sample_clocks) else None
        # ---- counts of this rank's share (identical every step; counted once, outside the timed region) ----
        if kind == "lines" and peer is not None:
            compute()                    # this rank's lines once more into its local buffer
            torch.cuda.synchronize()
        r2 = d_rec.reshape(-1, A.REC_DOUBLES)
        conv = ((r2[:, A.REC_STATUS].to(torch.int64) & 1) != 0).sum().to(torch.float64)
        # quadrature nodes a pass actually sweeps at each point: p_num when xi == 0 (isotropic collapse: the cos(theta) sum
        # is pre-summed into the weights), p_num * t_num otherwise — the algorithmic FLOP count follows the evaluated nodes
        nodes_pt = torch.where(r2[:, A.REC_XI] == 0.0, float(p if eng.isotropic_collapse else n_nodes), float(n_nodes))
        unit = lambda col: (r2[:, col] * nodes_pt).sum()
        cnt = torch.stack([conv, r2[:, A.REC_NEVAL].sum(), r2[:, A.REC_NTHERMO].sum(), r2[:, A.REC_NFUSED].sum(),
                           unit(A.REC_NEVAL), unit(A.REC_NTHERMO), unit(A.REC_NFUSED)])
        tot = torch.tensor([sum(step_ms), sum(kern_ms)], dtype=torch.float64, device=dev)
        mine_ms = torch.tensor([sum(kern_ms) / max(1, steps)], dtype=torch.float64, device=dev)
        local_sum = checksum(r2)
        all_ms = [torch.zeros_like(mine_ms) for _ in range(world)]
        # ---- the gathered array on rank 0 against the per-rank results ----
        verified, equal_nccl = None, None
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            sums = torch.stack([conv.to(torch.int64), local_sum])
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            dist.all_gather(all_ms, mine_ms)
            if kind == "lines":
                gathered = peer.tensor if peer is not None else full_holder[0]
                if rank == 0:
                    g2 = gathered.reshape(-1, A.REC_DOUBLES)
                    gconv = ((g2[:, A.REC_STATUS].to(torch.int64) & 1) != 0).sum()
                    verified = bool(gconv == sums[0]) and bool(checksum(g2) == sums[1])
                if peer is not None:
                    # once per run: the peer-stored array against a plain NCCL gather of the same records (bit for bit)
                    full, _ = scan_sharded(grid, len(xis), n_mu, lambda _l: d_rec, rank, world, layout=args.layout)
                    if rank == 0:
                        equal_nccl = bool(torch.equal(full.view(torch.int64), peer.tensor.view(torch.int64)))
                    del full
            elif rank == 0:
                g2 = torch.cat(gather_bufs).reshape(-1, A.REC_DOUBLES)
                gconv = ((g2[:, A.REC_STATUS].to(torch.int64) & 1) != 0).sum()
                verified = bool(gconv == sums[0]) and bool(checksum(g2) == sums[1])
        else:
            all_ms = [mine_ms]
        J.update(total_ms=float(tot[0]), kernel_ms=float(tot[1]), cnt=[float(v) for v in cnt], n_total=int(n_total),
                 per_rank_kernel_ms=[float(v[0]) for v in all_ms], gather_verified=verified, gather_equals_nccl=equal_nccl,
                 peer=peer is not None, stats=eng.stats(), n_mu=n_mu, n_T=n_T)
        if kind == "lines":
            J["host_lines"] = (grid, mine)
        if peer is not None:
            barrier()
            peer.close()
        WORKLOADS[w] = base_workload
        return J

    mu_factor = world if (args.scaling == "weak" and world > 1) else 1
    J = run_job(mu_factor, args.steps, args.warmup, True)
    total_ms, kernel_ms = J["total_ms"], J["kernel_ms"]
    n_conv, n_fj, n_th, n_ft, u_fj, u_th, u_ft = J["cnt"]
    n_total = J["n_total"]
    value = n_conv * args.steps / (total_ms * 1e-3)
    fl = alg_flops(1, u_fj, u_th, 2, u_ft)            # whole job, one step: evaluated nodes x flavours actually evaluated (u == d)
    fl_ref = alg_flops(n_nodes, n_fj, n_th, 3, n_ft)  # same passes counted the way the reference loops (full mesh, 3 flavours)
    st = J["stats"]

    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak and kind == "lines":
        Jw = run_job(world, min(args.steps, 2), 1, False)
        weak = {"value": Jw["cnt"][0] * min(args.steps, 2) / (Jw["total_ms"] * 1e-3), "unit": "points/s",
                "ms_per_step": Jw["total_ms"] / min(args.steps, 2), "points": Jw["n_total"],
                "note": "mu axis refined %d-fold: every GPU carries a full BASELINE-sized share" % world,
                "per_rank_kernel_ms": Jw["per_rank_kernel_ms"]}

    # ---- e2e: the reference-facing call with HOST buffers (pnjl_scan_lines_host / pnjl_solve_points_host):
    # H2D of the inputs, kernel, D2H of the records inside the timed region; pinned host memory.  With a page-locked
    # result buffer the kernels write the records straight into host memory while they run (include/pnjl_b200.h), so the
    # D2H bytes cross PCIe during the kernel instead of in a copy after it; pageable buffers would be staged and copied.
    e2e = None
    if not args.no_e2e:
        if kind == "lines":
            grid, mine = J["host_lines"]
            h_muq = torch.as_tensor(grid.muq_MeV[mine]).pin_memory().numpy()
            h_xi = torch.as_tensor(grid.xi[mine]).pin_memory().numpy()
            h_T = torch.as_tensor(grid.T_MeV).pin_memory().numpy()
            h_tidx = grid.table_idx[mine]
            h_rec = torch.empty((len(mine), J["n_T"], A.REC_DOUBLES), dtype=torch.float64).pin_memory().numpy()
            call = lambda: eng.scan_lines(h_muq, h_xi, h_T, h_tidx, out=h_rec)
            h2d = h_muq.nbytes + h_xi.nbytes + h_T.nbytes + h_tidx.nbytes
        else:
            allp, idx = J["host_inputs"]
            h_T = torch.as_tensor(allp[0, idx].copy()).pin_memory().numpy()
            h_mu = torch.as_tensor(allp[1, idx].copy()).pin_memory().numpy()
            h_xi = torch.as_tensor(allp[2, idx].copy()).pin_memory().numpy()
            h_rec = torch.empty((len(idx), A.REC_DOUBLES), dtype=torch.float64).pin_memory().numpy()
            call = lambda: eng.solve_points(h_T, h_mu, h_xi, A.SEED_MULTI, out=h_rec)
            h2d = h_T.nbytes + h_mu.nbytes + h_xi.nbytes
        call()                                          # warm-up (buffers inside the handle are grown here)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        hc = torch.tensor([float(((h_rec.reshape(-1, A.REC_DOUBLES)[:, A.REC_STATUS].astype(np.int64) & 1) != 0).sum())],
                          dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(hc, op=dist.ReduceOp.SUM)
        e2e = {"value": float(hc[0]) * args.steps / float(dt[0]), "unit": "points/s",
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(h_rec.nbytes) * world,
               "api": "pnjl_scan_lines_host" if kind == "lines" else "pnjl_solve_points_host",
               "d2h": "records written in place into the caller's page-locked buffer by the kernel (zero-copy over PCIe)"
                      if os.environ.get("PNJL_ZERO_COPY", "1") != "0" else "staged in HBM, cudaMemcpyAsync after the kernel",
               "ms_per_step": 1e3 * float(dt[0]) / args.steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = cpu_sample(w, target_seconds=args.cpu_seconds, engine="ad", step=1)
        sa = cpu_sample(w, target_seconds=min(args.cpu_seconds, 6.0), engine="analytic")
        cpu = {"value": s["converged"] / s["seconds"], "unit": "points/s", "cores": s["threads"], "kind": "port",
               "sample": s["sample"], "seconds": s["seconds"], "evals_per_point": s["n_fj"] / max(1, s["points"]),
               "note": "oracle/pnjl_oracle.cpp: the reference's arithmetic (nested-dual AD Jacobian like ForwardDiff inside NLsolve)",
               "analytic_jacobian": {"value": sa["converged"] / sa["seconds"], "unit": "points/s", "cores": sa["threads"],
                                     "kind": "port", "sample": sa["sample"], "seconds": sa["seconds"],
                                     "evals_per_point": sa["n_fj"] / max(1, sa["points"]),
                                     "note": "oracle/pnjl_analytic_cpu.cpp: the GPU kernel's algorithm (closed-form Jacobian, isospin "
                                             "shortcut, fused final pass) compiled for the host, OpenMP over lines (SURVEY.md §8d)"}}

    if rank == 0:
        traffic = None
        pj = os.path.join(ROOT, "profiles", "top_kernel.json")
        static = None
        if os.path.exists(pj) and world == 1 and w == "cfg5":      # the committed capture is the cfg5 launch of k_solve_ws
            try:
                prof = json.load(open(pj))
                if prof.get("dram_bytes_per_point") is not None:
                    traffic = prof["dram_bytes_per_point"] * float(n_total)          # per launch
                static = {k: prof[k] for k in ("fp64_pipe_util_pct", "issue_active_pct", "profile", "commit", "kernel") if k in prof}
                static["source"] = "static: ncu capture committed under profiles/ (not measured in this run)"
            except Exception:
                pass
        achieved = fl * args.steps / (kernel_ms * 1e-3) / 1e12
        per_rank = J["per_rank_kernel_ms"]
        out = {
            "metric": "converged PNJL gap points/sec", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, world, args.scaling, args.layout, {
                "points": int(n_total), "converged": int(n_conv), "n_mu": int(J["n_mu"]),
                "sharding": "mu dealt out %s over %d rank(s)" % ("round-robin" if args.layout == "interleaved" else "in contiguous slabs", world) + (
                    "" if world == 1 else ("; records stored by every rank's kernel directly into rank 0's array (CUDA IPC, NVLink), "
                                           "closing barrier only" if J["peer"] else "; dist.gather to rank 0 after the kernel")),
                "l2": "512 MB buffer written between timed steps (L2 flush); inputs are O(100 KB), outputs 256 B/point",
                "lanes_per_solve": st["lanes_per_solve"], "blocks": st["blocks"], "threads": st["threads"],
                "regs_per_thread": st["regs_per_thread"]}),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_sus * world, "unit": "TFLOP/s",
                         "frac": achieved / (peak_sus * world), "traffic": traffic,
                         "peak_source": "DFMA microbenchmark run in this process (pnjl_measure_fp64_peak): sustained %.2f, "
                                        "burst %.2f TFLOP/s per GPU; MEASURED_PEAKS.json has no FP64 figure" % (peak_sus, peak_burst),
                         "algorithmic_flop_per_step": fl, "flavours_evaluated": 2,
                         "node_evaluations_per_step": u_fj + u_th + u_ft,
                         "isotropic_collapse": bool(eng.isotropic_collapse),
                         "reference_equivalent_tflops": fl_ref * args.steps / (kernel_ms * 1e-3) / 1e12,
                         "fj_passes_per_point": n_fj / max(1.0, float(n_total)),
                         "thermo_passes_per_point": n_th / max(1.0, float(n_total)),
                         "fused_final_passes_per_point": n_ft / max(1.0, float(n_total)),
                         "passes_per_point": (n_fj + n_th + n_ft) / max(1.0, float(n_total)),
                         "kernel_ms_per_step": kernel_ms / args.steps,
                         "static_profile": static},
            "per_rank_kernel_ms": per_rank,
            "rank_imbalance_max_over_mean": max(per_rank) / (sum(per_rank) / len(per_rank)),
            "gather_verified": J["gather_verified"], "gather_equals_nccl": J["gather_equals_nccl"],
            "gpu_launches": int(args.steps) * world * int(st["kernel_launches"]),
            "clocks": J["clocks"],
        }
        if weak is not None:
            out["weak"] = weak
        if e2e is not None:
            out["e2e"] = e2e
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
