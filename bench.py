#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 sparse direct solver backend.

Metric (BASELINE.json): FP64 factor GFLOP/s (and factor+solve seconds) on the 3-D 7-point
Poisson 128^3 SPD system.  One "step" = factorize + evaluate(p=1) of that system.

  value      F / (mean factorize seconds), F = sum_j colcount_j^2 from the analysis' exact column
             counts; values already resident in HBM (b200_factor_device), timed with CUDA events
  e2e        the same metric through the reference-facing plugin callbacks
             (solver_factorize / solver_evaluate behind solver_lookup[]) with HOST buffers:
             F / (wall seconds of the factorize tick, H2D of A->dd inside)
  roofline   the DMMA trailing-update kernel (gemm_nt_dmma): algorithmic flops / its launch time
  cpu_baseline  the oracle's BLAS-3 multifrontal Cholesky on the host cores, bounded sample

`--impl reference` times the CPU implementation of the path (oracle port; the reference's own
CHOLMOD/UMFPACK/MUMPS cannot be built in this image) on the same metric.
N > 1 (torchrun, one rank per GPU): ONE factorization is spread over the N GPUs -- independent
elimination-tree subtrees go to single ranks, the top separator fronts are distributed block-cyclically
by column over rank groups, panels / contribution-block columns / finished factor panels move with NCCL
over NVLink (csrc/host/b200_plan.c).  The problem size is fixed -> "scaling": "strong";
value = F / (max over ranks of the device-timed factor seconds).
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_pkg():
    d = os.path.join(ROOT, "meagre-crowd_b200")
    spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules["meagre_crowd_b200"] = m
    spec.loader.exec_module(m)
    return m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu = gpu; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for l in self.proc.stdout:
            self.lines.append((time.perf_counter(), l))

    def stop(self, t_from=None):
        """samples taken since t_from (the start of the timed region); the sampler itself is started before the warm-up, because nvidia-smi needs
        longer to come up than a short timed region lasts -- if the region saw no sample, the samples of the warm-up steps (the same load) are used and the line says so"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [l for t, l in self.lines if t_from is None or t >= t_from]
        window = "timed region"
        if not inside:
            inside = [l for t, l in self.lines][-20:]; window = "warm-up steps (no sample fell into the timed region)"
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def configs(stencils, m):
    """BASELINE.json configs[1..4] (configs[0] is the 5x5 plumbing fixture, covered by the tests)"""
    return {
        "c2": dict(gen=lambda: stencils.poisson2d_upper(1024), kind=m.KIND_CHOLESKY, sym=True, workload="poisson2d_5pt_1024^2_spd_cholesky"),
        "c3": dict(gen=lambda: stencils.poisson3d_upper(128), kind=m.KIND_CHOLESKY, sym=True, workload="poisson3d_7pt_128^3_spd_cholesky"),
        "c4": dict(gen=lambda: stencils.convdiff27_full(96), kind=m.KIND_LU, sym=False, workload="convdiff3d_27pt_96^3_unsymmetric_lu"),
        "c5": dict(gen=lambda: stencils.poisson3d_upper(128), kind=m.KIND_CHOLESKY, sym=True, workload="poisson3d_7pt_128^3_spd_cholesky+multi_rhs_sweep"),
    }


def scaled_berr(A, x, b):
    """||Ax-b|| / (||A|| ||x|| + ||b||), infinity norms (north_star)"""
    import scipy.sparse.linalg as sla
    return float(np.abs(A @ x - b).max() / (sla.norm(A, np.inf) * np.abs(x).max() + np.abs(b).max()))


def rhs_sweep(m, s, n, nnzL, hbm_peak, fp64_peak, ps=(1, 2, 4, 8, 16, 32, 64, 128, 256)):
    """C5: solve seconds for p right-hand sides on the factor held by `s` (b and x resident in HBM, no refinement),
    against the roof that bounds it: HBM (two sweeps over L) below the crossover, FP64 above"""
    L = s.L; ctx = s.ctx; st = m.b200_stats_t()
    L.b200_set_refinement(ctx, 0)
    out = {}
    for p in ps:
        B = np.asfortranarray(((np.arange(n, dtype=np.int64)[:, None] + np.arange(p, dtype=np.int64)[None, :]) % 1024).astype(np.float64))
        d_b = L.b200_dev_alloc(ctx, B.nbytes); d_x = L.b200_dev_alloc(ctx, B.nbytes)
        if not d_b or not d_x:
            break
        L.b200_dev_upload(ctx, d_b, B.ctypes.data, B.nbytes)
        best = 1e30
        for _ in range(3):
            if L.b200_solve_device(ctx, d_b, p, d_x) != 0:
                raise RuntimeError("solve failed: " + s.err())
            L.b200_get_stats(ctx, C.byref(st)); best = min(best, st.solve_ms)
        t = best * 1e-3
        gbs = (16.0 * nnzL + 32.0 * n * p) / t / 1e9; tfs = 4.0 * nnzL * p / t / 1e12
        out[str(p)] = {"solve_s": t, "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "fp64_tflops": tfs, "fp64_frac": tfs / fp64_peak}
        L.b200_dev_free(ctx, d_b); L.b200_dev_free(ctx, d_x)
    L.b200_set_refinement(ctx, -1)      # back to automatic
    return out


def side_config(m, cfg, hbm_peak, fp64_peak):
    """one of the other BASELINE configs at full size on this GPU: factor / solve seconds (device events, values resident in HBM),
    backward error, fractions of the FP64 roof (factor) and of the HBM roof (solve)"""
    import ctypes as C_
    n, cp, ri, v = cfg["gen"]()
    LU = cfg["kind"] == m.KIND_LU
    s = m.Solver(device=0); L = s.L
    t0 = time.perf_counter(); S = s.analyze(n, cp, ri, cfg["kind"]); t_an = time.perf_counter() - t0
    F = S.flops * (2.0 if LU else 1.0); nnzL = int(S.nnzL)
    d_vals = L.b200_dev_alloc(s.ctx, v.nbytes); L.b200_dev_upload(s.ctx, d_vals, v.ctypes.data, v.nbytes)
    b = np.arange(n, dtype=np.float64)
    d_b = L.b200_dev_alloc(s.ctx, b.nbytes); d_x = L.b200_dev_alloc(s.ctx, b.nbytes); L.b200_dev_upload(s.ctx, d_b, b.ctypes.data, b.nbytes)
    st = m.b200_stats_t(); fms, sms = [], []
    for i in range(4):
        if L.b200_factor_device(s.ctx, d_vals) != 0:
            raise RuntimeError("factor failed: " + s.err())
        L.b200_get_stats(s.ctx, C_.byref(st)); f_ = st.factor_ms; pert = st.npiv_perturbed
        if L.b200_solve_device(s.ctx, d_b, 1, d_x) != 0:
            raise RuntimeError("solve failed: " + s.err())
        L.b200_get_stats(s.ctx, C_.byref(st))
        if i >= 1:
            fms.append(f_); sms.append(st.solve_ms)
    x = np.empty(n); L.b200_dev_download(s.ctx, x.ctypes.data, d_x, x.nbytes)
    A = stencils_mod().to_scipy(n, cp, ri, v, cfg["sym"])
    fs = float(np.mean(fms)) * 1e-3; ss = float(np.mean(sms)) * 1e-3
    sol_bytes = (16.0 * nnzL + 32.0 * n) * (1 + st.refine_steps)     # L once forward, L' (or U') once backward, x read+written per sweep
    out = {"workload": cfg["workload"], "n": n, "nnz_L": nnzL, "factor_flops": F, "analyze_s": t_an, "factor_s": fs, "solve_s": ss, "factor_plus_solve_s": fs + ss,
           "factor_tflops": F / fs / 1e12, "factor_frac_of_fp64_roof": F / fs / 1e12 / fp64_peak, "solve_hbm_gbs": sol_bytes / ss / 1e9, "solve_frac_of_hbm_roof": sol_bytes / ss / 1e9 / hbm_peak,
           "refine_steps": int(st.refine_steps), "npiv_perturbed": int(pert), "backward_error": scaled_berr(A, x, b), "supernodes": int(S.ns), "max_front": int(S.maxfront), "device_gb": st.device_bytes / 1e9}
    L.b200_dev_free(s.ctx, d_vals); L.b200_dev_free(s.ctx, d_b); L.b200_dev_free(s.ctx, d_x)
    s.close()
    return out


def stencils_mod():
    from meagre_crowd_b200 import stencils
    return stencils


def cpu_factor_sample(m, grid, threads=0, reps=1):
    """time the oracle's CPU multifrontal Cholesky (BLAS-3 via the bundled OpenBLAS) on a poisson3d grid^3 sample"""
    import oracle
    from meagre_crowd_b200 import stencils
    n, cp, ri, v = stencils.poisson3d_upper(grid)
    S = m.analyze(n, cp, ri)
    s = S.contents
    ns = s.ns
    sym = dict(ns=ns, n=n, maxdepth=s.maxdepth)
    for name, ln, dt in (("perm", n, np.int32), ("iperm", n, np.int32), ("sn_start", ns + 1, np.int32), ("sn_parent", ns, np.int32),
                         ("sn_depth", ns, np.int32), ("rowptr", ns + 1, np.int64), ("lptr", ns + 1, np.int64)):
        sym[name] = m.sym_array(S, name, ln, dt)
    sym["rowidx"] = m.sym_array(S, "rowidx", int(sym["rowptr"][-1]), np.int32)
    flops = s.flops
    has_blas = oracle.load_blas()
    cores = threads or os.cpu_count()
    best = 1e30; times = []
    for _ in range(reps):
        t = time.perf_counter(); panels = oracle.mf_cholesky(n, cp, ri, v, sym, threads=cores); dt = time.perf_counter() - t
        times.append(dt); best = min(best, dt)
    b = np.arange(n, dtype=np.float64)
    t = time.perf_counter(); x = oracle.mf_solve(sym, panels, b); tsolve = time.perf_counter() - t
    A = stencils.to_scipy(n, cp, ri, v, True)
    berr = float(np.abs(A @ x - b).max() / (6.0 * 2 * np.abs(x).max() + np.abs(b).max()))
    m.lib().b200_symbolic_free(S)
    return dict(flops=flops, times=times, best=best, solve_s=tsolve, cores=cores, blas=has_blas, berr=berr, n=n)


def run_reference(args, m, rank, world):
    """--impl reference: the CPU implementation of the path on the box's host cores (rank 0 only).
    The reference's own numerics (CHOLMOD / MUMPS) are not in the image, so this is the oracle's BLAS-3 multifrontal
    Cholesky (OpenBLAS, all host cores) -- on the SAME system as the b200 arm (poisson3d 128^3) when the host has the
    memory for it (~35 GB), with the step loop capped so that the run ends within minutes; else on the largest grid that fits.
    It borrows the repo's ordering / symbolic analysis (libmeagre_crowd_b200.so is loaded for that and for nothing else)."""
    if rank != 0:
        return
    try:
        import psutil
        avail = psutil.virtual_memory().available / 1e9
    except Exception:
        avail = 0.0
    grid = args.cpu_grid if args.cpu_grid_given else (128 if avail >= 48.0 else 96 if avail >= 16.0 else 64)
    # budget: F(128^3) = 1.15e13 flop at 50-100 GFLOP/s is 2-4 minutes per factorization -> one untimed small warm-up, capped timed steps
    cpu_factor_sample(m, 32)
    est = {128: 200.0, 96: 40.0}.get(grid, 4.0)
    steps = max(1, min(args.steps, int(240.0 / est)))
    r = cpu_factor_sample(m, grid, reps=steps)
    mean = float(np.mean(r["times"]))
    val = r["flops"] / mean / 1e9
    line = {"impl": "reference", "metric": "fp64_factor_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
            "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"poisson3d_7pt_{grid}^3_spd_cholesky",
                       "step": f"CPU multifrontal Cholesky factorize of poisson3d {grid}^3 on the host cores (oracle port of the path: BLAS-3 on dense fronts via the bundled OpenBLAS; "
                               f"the repo's own nested-dissection ordering and symbolic analysis are borrowed, so both arms factor the same tree); step loop capped at {steps} of the requested {args.steps}, warm-up on a 32^3 grid",
                       "requested_steps": args.steps, "requested_warmup": args.warmup, "host_mem_available_gb": avail},
            "factor_s": mean, "solve_s": r["solve_s"], "factor_plus_solve_s": mean + r["solve_s"], "backward_error": r["berr"], "factor_flops": r["flops"],
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": r["cores"], "kind": "port",
                             "sample": f"poisson3d 7-pt {grid}^3 (n={r['n']}, F={r['flops']:.3e} flop), factorize {mean:.1f}s + solve {r['solve_s']:.2f}s, BLAS-3 via bundled OpenBLAS={r['blas']}, berr={r['berr']:.1e}"},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c5"], help="BASELINE.json configs[1..4]; c3 = the headline (default), c5 = c3 + the multi-RHS sweep")
    ap.add_argument("--grid", type=int, default=0, help="override the grid edge of the chosen config's stencil (0 = the BASELINE size)")
    ap.add_argument("--no-side-configs", action="store_true", help="headline run only: skip the short c2 / c4 / c5 legs reported under `configs`")
    ap.add_argument("--cpu-grid", type=int, default=0, help="grid edge of the CPU leg (default: 80 for the cpu_baseline sample of the b200 arm; 128 for --impl reference when the host has the memory)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.args_grid = args.grid
    args.cpu_grid_given = args.cpu_grid > 0
    if not args.cpu_grid_given and args.impl == "b200":
        args.cpu_grid = 80
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    m = load_pkg()
    if not os.path.exists(m.LIB_PATH):
        m.build()
    if args.impl == "reference":
        run_reference(args, m, rank, world)
        return
    from meagre_crowd_b200 import stencils
    L = m.lib()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- problem + analysis (not timed: "analyze" is a separate tick in the reference, src/solvers.c:478)
    cfgs = configs(stencils, m); cfg = cfgs[args.config]
    KIND = cfg["kind"]; LU = KIND == m.KIND_LU
    if args.grid:
        gen = {"c2": lambda: stencils.poisson2d_upper(args.grid), "c4": lambda: stencils.convdiff27_full(args.grid)}.get(args.config, lambda: stencils.poisson3d_upper(args.grid))
        workload = cfg["workload"].replace("1024^2", f"{args.grid}^2").replace("128^3", f"{args.grid}^3").replace("96^3", f"{args.grid}^3")
    else:
        gen = cfg["gen"]; workload = cfg["workload"]
    if world > 1 and LU:
        raise SystemExit("the multi-GPU factorization is built for Cholesky (c2/c3/c5); run c4 on one GPU")
    n, cp, ri, v = gen()
    s = m.Solver(device=local)
    if world > 1:
        s.join(rank, world, dist=dist)
    host_threads = max(1, (os.cpu_count() or 8) // world)
    # rank 0 orders with all host cores (N > 1: and broadcasts the permutation); every rank then runs the symbolic analysis on the GIVEN permutation.
    # N = 1 takes the same two steps, so that every N factors the very same supernodal tree and the checksums below can be compared across the N lines
    # (the analysis is not idempotent on its own output permutation: postorder and amalgamation renumber the columns).  analyze_s is the first step,
    # which is what the plugin's solver_analyze does; the second step is reported as reanalyze_s.
    t0 = time.perf_counter()
    perm = m.shared_ordering(n, cp, ri, KIND, rank, world, dist, threads=os.cpu_count() or 8)
    t_analyze = time.perf_counter() - t0
    t0 = time.perf_counter()
    S = s.analyze(n, cp, ri, KIND, perm=perm, threads=host_threads)
    t_reanalyze = time.perf_counter() - t0
    F = S.flops * (2.0 if LU else 1.0)
    ctx = s.ctx
    d_vals = L.b200_dev_alloc(ctx, v.nbytes); L.b200_dev_upload(ctx, d_vals, v.ctypes.data, v.nbytes)
    b = np.arange(n, dtype=np.float64)
    d_b = L.b200_dev_alloc(ctx, b.nbytes); d_x = L.b200_dev_alloc(ctx, b.nbytes); L.b200_dev_upload(ctx, d_b, b.ctypes.data, b.nbytes)
    st = m.b200_stats_t()
    refine = [0]

    def step():
        r = L.b200_factor_device(ctx, d_vals)
        if r != 0:
            raise RuntimeError("factor failed: " + s.err())
        L.b200_get_stats(ctx, C.byref(st)); fms = st.factor_ms; nl = st.launches
        if rank != 0:       # N > 1: the finished factor is collected on rank 0, which runs the sweeps
            return fms, 0.0, nl
        r = L.b200_solve_device(ctx, d_b, 1, d_x)
        if r != 0:
            raise RuntimeError("solve failed: " + s.err())
        L.b200_get_stats(ctx, C.byref(st))
        refine[0] = st.refine_steps
        return fms, st.solve_ms, nl + st.solve_launches

    def barrier():
        L.b200_dev_sync(ctx)
        if dist is not None:
            import torch
            dist.barrier(); torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    t0 = time.perf_counter()
    fms, sms, launches = [], [], 0
    for _ in range(args.steps):
        a, b_, nl = step(); fms.append(a); sms.append(b_); launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop(t0)
    factor_s = float(np.mean(fms)) * 1e-3; solve_s = float(np.mean(sms)) * 1e-3
    # max over ranks of the device-timed factor and of the region
    if dist is not None:
        import torch
        t = torch.tensor([factor_s, solve_s, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        factor_s, solve_s, wall = t.tolist()
        tl = torch.tensor([launches], device="cuda", dtype=torch.int64); dist.all_reduce(tl); launches = int(tl.item())
    value = F / factor_s / 1e9     # one factorization shared by all ranks (strong scaling)
    L.b200_get_stats(ctx, C.byref(st)); nvlink_bytes = float(st.comm_bytes)
    if dist is not None:
        import torch
        tb_ = torch.tensor([nvlink_bytes], device="cuda", dtype=torch.float64); dist.all_reduce(tb_); nvlink_bytes = tb_.item() / 2.0   # every message is counted by its sender and its receiver

    # correctness gate on the timed configuration: scaled backward error <= 1e-12
    x = np.zeros(n)
    if rank == 0:
        L.b200_dev_download(ctx, x.ctypes.data, d_x, x.nbytes)
    A = stencils.to_scipy(n, cp, ri, v, cfg["sym"])
    bvec = np.arange(n, dtype=np.float64)
    berr = scaled_berr(A, x, bvec) if rank == 0 else 0.0
    # checksums that must not depend on the number of GPUs: ||L||_F^2 (Cholesky: = trace(A)) and a hash of the bits of x
    import hashlib
    x_sha = hashlib.sha256(x.tobytes()).hexdigest()[:16]
    colsq = np.empty(n); L.b200_factor_colnorms(ctx, colsq.ctypes.data)
    if dist is not None:
        import torch
        tcs = torch.from_numpy(colsq).cuda(); dist.all_reduce(tcs, op=dist.ReduceOp.MAX); colsq = tcs.cpu().numpy()      # a column's owner holds its sum, the others 0 (or the same bits)
    l_frob2 = float(np.sum(colsq)); l_sha = hashlib.sha256(colsq.tobytes()).hexdigest()[:16]

    roofline = None; cpu_base = None; e2e = None; extra = {}
    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream (profile mode serialises the launches;
    # with N > 1 the factorization is collective, so every rank runs this pass and rank 0 reports its own kernels)
    L.b200_set_profile(ctx, 1)
    L.b200_factor_device(ctx, d_vals); L.b200_get_stats(ctx, C.byref(st))
    L.b200_set_profile(ctx, 0)
    if rank == 0:
        buf = C.create_string_buffer(16384); L.b200_get_profile(ctx, buf, 16384)
        prof = {l.split()[0]: (int(l.split()[1].split("=")[1]), float(l.split()[2].split("=")[1])) for l in buf.value.decode().strip().splitlines() if " launches=" in l}
        gemm_ms = st.gemm_ms; gemm_alg = st.gemm_alg_flops
        dmma_peak = L.b200_bench_fp64_peak_tflops(ctx, 0, 3); dfma_peak = L.b200_bench_fp64_peak_tflops(ctx, 1, 3)
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); hbm_measured = True
        except Exception:
            hbm_peak = 6650.0; hbm_measured = False
        peak = 37.0   # nominal HGX B200 FP64 (tensor = vector); MEASURED_PEAKS.json has no FP64 entry -- see DESIGN.md
        achieved = gemm_alg / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        roofline = {"bound": "tensor", "kernel": "gemm_nt_dmma_bulk_kernel (DMMA.8x8x4 via mma.sync m8n8k4.f64 -- tcgen05 has no f64 kind; 64x64 CTA tile, cp.async.bulk + mbarrier producer/consumer pipeline, 3 CTAs/SM)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if achieved else None, "traffic": 1.122e9,
                    "traffic_note": "dram__bytes_read+write of ONE captured launch at a root-front shape (8192x8192, K=512: algorithmic 1.14e9 B, DMMA pipe 89.1 % active, 32.8 TFLOP/s) -- profiles/r1_ncu_prof_gemm_root.txt; `achieved` is the average over all update launches of the factorization",
                    "peak_source": "nominal FP64 37 TFLOP/s (HGX B200 datasheet; not in MEASURED_PEAKS.json)",
                    "measured_dmma_issue_roof_tflops": dmma_peak, "measured_dfma_roof_tflops": dfma_peak,
                    "frac_of_measured_dmma_roof": achieved / dmma_peak if achieved and dmma_peak > 0 else None,
                    "gemm_ms_per_factor": gemm_ms, "gemm_launches": int(st.gemm_launches), "gemm_share_of_factor": gemm_ms / sum(x[1] for x in prof.values()) if prof else None,
                    "kernel_ms": {k: v_[1] for k, v_ in prof.items()}}
        ns_ = S.ns
        k_ = np.diff(m.sym_array(s.S, "sn_start", ns_ + 1, np.int32)).astype(np.float64); f_ = np.diff(m.sym_array(s.S, "rowptr", ns_ + 1, np.int64)).astype(np.float64)
        u_ = f_ - k_; ea_bytes = float(16.0 * (u_ * (u_ + 1) / 2).sum())
        if world > 1:
            prof["extend_add"] = (0, 0.0)      # whole-tree bytes over one rank's time is not a rate: reported at N=1 only
        extra = {"factor_s": factor_s, "solve_s": solve_s, "factor_plus_solve_s": factor_s + solve_s, "analyze_s": t_analyze, "reanalyze_s": t_reanalyze, "backward_error": berr,
                 "n": n, "nnz_L": int(S.nnzL), "stored_nnz_L": int(S.stored_nnzL), "factor_flops": F, "supernodes": S.ns, "levels": S.maxdepth + 1, "max_front": S.maxfront,
                 "device_gb": st.device_bytes / 1e9, "solve_hbm_gbs": (16.0 * S.nnzL + 32.0 * n) * (1 + refine[0]) / solve_s / 1e9, "refine_steps": refine[0], "nccl_bytes_per_factor": nvlink_bytes,
                 "extend_add_roofline": {"bound": "hbm", "kernel": "extend_add_kernel", "achieved": ea_bytes / (prof["extend_add"][1] * 1e-3) / 1e9 if prof.get("extend_add", (0, 0))[1] > 0 else None,
                                         "peak": hbm_peak, "unit": "GB/s", "frac": ea_bytes / (prof["extend_add"][1] * 1e-3) / 1e9 / hbm_peak if prof.get("extend_add", (0, 0))[1] > 0 else None,
                                         "note": "algorithmic bytes = 16 B x sum_s u_s(u_s+1)/2 (read the child's entry, write the parent's; the read of the parent entry is not counted); ncu at the root level: DRAM traffic 8.7 GB = read child + read/write parent, 3.2 TB/s (profiles/r1_ncu_prof_extend_add.txt)"},
                 "solve_roofline": {"bound": "hbm", "kernels": "solve_fwd_step_kernel / solve_bwd_step_kernel (two sweeps over the factor)", "achieved": (16.0 * S.nnzL + 32.0 * n) * (1 + refine[0]) / solve_s / 1e9,
                                    "peak": hbm_peak, "unit": "GB/s", "frac": (16.0 * S.nnzL + 32.0 * n) * (1 + refine[0]) / solve_s / 1e9 / hbm_peak,
                                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if hbm_measured else "fallback 6650 GB/s (B200_PROFILING.md)",
                                    "note": "algorithmic bytes = 16 B x exact nnz(L) + 32 B x n per solve (SURVEY 8d): L streamed once per sweep, x read+written per sweep; relaxation zeros are not counted"}}
    sweep = None
    if rank == 0 and world == 1 and (args.config == "c5" or (args.config == "c3" and not args.args_grid and not args.no_side_configs)):
        sweep = rhs_sweep(m, s, n, int(S.nnzL), hbm_peak, 37.0)     # C5 on the factor that is already resident
    # release the device-resident copy before the plugin path allocates its own
    L.b200_dev_free(ctx, d_vals); L.b200_dev_free(ctx, d_b); L.b200_dev_free(ctx, d_x)
    s.close()
    side = {}
    if rank == 0 and world == 1 and args.config == "c3" and not args.args_grid and not args.no_side_configs:
        for key in ("c2", "c4"):
            try:
                side[key] = side_config(m, cfgs[key], hbm_peak, 37.0)
            except Exception as e:      # a side leg must never take the headline line down
                side[key] = {"error": str(e)}
    if rank == 0:
        extra["checksums"] = {"x_sha256_16": x_sha, "L_frob2": l_frob2, "L_colnorms_sha256_16": l_sha,
                              "note": "bit-identical factor and solution on any number of GPUs: these three values must be equal in the N=1,2,4,8 lines"}
        if sweep is not None:
            side["c5"] = {"workload": cfgs["c5"]["workload"], "rhs": "B[i,k] = (i+k) mod 1024, resident in HBM, no refinement, best of 3", "solve_by_p": sweep}
        if side:
            extra["configs"] = side

    if world > 1 and not args.no_e2e:
        # ---- e2e at N > 1: the C-ABI shim with HOST buffers on every rank (H2D of the values / b and D2H of x inside the timed calls);
        # the plugin row itself is single-process (SOLVER_SINGLE_THREADED_ONLY), see DESIGN.md 7
        s2 = m.Solver(device=local); s2.join(rank, world, dist=dist)
        s2.analyze(n, cp, ri, KIND, perm=perm, threads=host_threads)
        ftimes, etimes = [], []
        for i in range(2 + 3):
            dist.barrier(); t = time.perf_counter(); rc = s2.factor(v); tf = time.perf_counter() - t
            t = time.perf_counter(); xe = s2.solve(bvec) if rank == 0 else bvec; te = time.perf_counter() - t
            if rc != 0:
                raise RuntimeError("factor failed: " + s2.err())
            if i >= 2:
                ftimes.append(tf); etimes.append(te)
        import torch
        tt = torch.tensor([float(np.mean(ftimes)), float(np.mean(etimes))], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        berr_e2e = scaled_berr(A, xe, bvec) if rank == 0 else 0.0
        s2.close()
        e2e = {"value": F / tt[0].item() / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(world * (v.nbytes + bvec.nbytes)), "d2h_bytes_per_step": int(world * bvec.nbytes),
               "factorize_s": tt[0].item(), "evaluate_s": tt[1].item(), "backward_error": berr_e2e,
               "path": "b200_factor_host / b200_solve_host (C-ABI shim, pageable host buffers) on every rank, max over ranks"}
    if world == 1 and rank == 0 and not args.no_e2e:
        # ---- e2e through the plugin table with host buffers (what `meagre-crowd -i A.mm -t -s b200` times)
        idx = L.lookup_solver_by_shortname(b"b200")
        A_m = m.make_matrix_csc(n, cp, ri, v, sym=m.SM_SYMMETRIC, location=m.UPPER_TRIANGULAR) if cfg["sym"] else m.make_matrix_csc(n, cp, ri, v)
        b_m = m.make_matrix_dcol(bvec); x_m = L.malloc_matrix()
        timer = L.perftimer_malloc()
        state = L.solver_init(idx, 0, 0, timer)
        L.solver_analyze(state, A_m)
        ftimes, etimes = [], []
        for i in range(2 + 3):
            t = time.perf_counter(); L.solver_factorize(state, A_m); tf = time.perf_counter() - t
            t = time.perf_counter(); L.solver_evaluate(state, b_m, x_m); te = time.perf_counter() - t
            if i >= 2:
                ftimes.append(tf); etimes.append(te)
        xe = m.matrix_to_numpy(x_m)[:, 0]
        berr_e2e = scaled_berr(A, xe, bvec)
        L.solver_finalize(state); L.free_matrix(A_m); L.free_matrix(b_m); L.free_matrix(x_m); L.perftimer_free(timer)
        e2e = {"value": F / float(np.mean(ftimes)) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(v.nbytes + bvec.nbytes), "d2h_bytes_per_step": int(bvec.nbytes),
               "factorize_s": float(np.mean(ftimes)), "evaluate_s": float(np.mean(etimes)), "backward_error": berr_e2e,
               "path": "solver_factorize/solver_evaluate via solver_lookup[] row 'b200', matrix_t in pageable host memory"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # the CPU baseline is reported at N=1 only
        r = cpu_factor_sample(m, args.cpu_grid)
        cpu_base = {"value": r["flops"] / r["best"] / 1e9, "unit": "GFLOP/s", "cores": r["cores"], "kind": "port",
                    "sample": f"poisson3d 7-pt {args.cpu_grid}^3 (n={r['n']}, F={r['flops']:.3e} flop), factorize {r['best']:.2f}s + solve {r['solve_s']:.2f}s, oracle multifrontal Cholesky, BLAS-3 via bundled OpenBLAS={r['blas']}, berr={r['berr']:.1e}"}
    if rank == 0:
        line = {"metric": "fp64_factor_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "step": "factorize + evaluate(p=1, automatic refinement), values and b resident in HBM",
                           "ordering": "nested dissection (own)", "l2": "inputs larger than L2 (factor touches >18 GB per step)",
                           "multi_gpu": f"one factorization over {world} GPUs: subtrees -> ranks, top fronts 1-D block-cyclic, NCCL panels/contribution blocks; finished panels collected on rank 0, which runs the sweeps" if world > 1 else "single GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
