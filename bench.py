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
            self.lines.append(l)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


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
    """--impl reference: the CPU implementation of the path on the box's host cores (rank 0 only)"""
    if rank != 0:
        return
    grid = args.cpu_grid
    for _ in range(max(args.warmup, 0)):
        cpu_factor_sample(m, min(grid, 32))
    r = cpu_factor_sample(m, grid, reps=args.steps)
    mean = float(np.mean(r["times"]))
    val = r["flops"] / mean / 1e9
    line = {"impl": "reference", "metric": "fp64_factor_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "poisson3d_7pt_128^3_spd_cholesky", "step": f"bounded sample: CPU multifrontal Cholesky factorize of poisson3d {grid}^3 (same generator, same ordering code)"},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": r["cores"], "kind": "port",
                             "sample": f"poisson3d 7-pt {grid}^3 (n={r['n']}, F={r['flops']:.3e} flop), factorize only, BLAS-3 via bundled OpenBLAS={r['blas']}, berr={r['berr']:.1e}"},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=128, help="poisson3d grid edge (headline: 128)")
    ap.add_argument("--cpu-grid", type=int, default=64, help="grid edge of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
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
    n, cp, ri, v = stencils.poisson3d_upper(args.grid)
    s = m.Solver(device=local)
    if world > 1:
        s.join(rank, world, dist=dist)
    host_threads = max(1, (os.cpu_count() or 8) // world)
    t0 = time.perf_counter()
    if world > 1:   # rank 0 orders with all host cores and broadcasts the permutation; every rank then runs the same symbolic analysis on it
        perm = m.shared_ordering(n, cp, ri, m.KIND_CHOLESKY, rank, world, dist, threads=os.cpu_count() or 8)
        S = s.analyze(n, cp, ri, m.KIND_CHOLESKY, perm=perm, threads=host_threads)
    else:
        perm = None
        S = s.analyze(n, cp, ri, m.KIND_CHOLESKY, threads=host_threads)
    t_analyze = time.perf_counter() - t0
    F = S.flops
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

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t0 = time.perf_counter()
    fms, sms, launches = [], [], 0
    for _ in range(args.steps):
        a, b_, nl = step(); fms.append(a); sms.append(b_); launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
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
    x = np.empty(n); L.b200_dev_download(ctx, x.ctypes.data, d_x, x.nbytes)
    A = stencils.to_scipy(n, cp, ri, v, True)
    bvec = np.arange(n, dtype=np.float64)
    berr = float(np.abs(A @ x - bvec).max() / (12.0 * np.abs(x).max() + np.abs(bvec).max()))

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
        extra = {"factor_s": factor_s, "solve_s": solve_s, "factor_plus_solve_s": factor_s + solve_s, "analyze_s": t_analyze, "backward_error": berr,
                 "n": n, "nnz_L": int(S.nnzL), "stored_nnz_L": int(S.stored_nnzL), "factor_flops": F, "supernodes": S.ns, "levels": S.maxdepth + 1, "max_front": S.maxfront,
                 "device_gb": st.device_bytes / 1e9, "solve_hbm_gbs": (16.0 * S.nnzL + 32.0 * n) * (1 + refine[0]) / solve_s / 1e9, "refine_steps": refine[0], "nccl_bytes_per_factor": nvlink_bytes,
                 "extend_add_roofline": {"bound": "hbm", "kernel": "extend_add_kernel", "achieved": ea_bytes / (prof["extend_add"][1] * 1e-3) / 1e9 if prof.get("extend_add", (0, 0))[1] > 0 else None,
                                         "peak": hbm_peak, "unit": "GB/s", "frac": ea_bytes / (prof["extend_add"][1] * 1e-3) / 1e9 / hbm_peak if prof.get("extend_add", (0, 0))[1] > 0 else None,
                                         "note": "algorithmic bytes = 16 B x sum_s u_s(u_s+1)/2 (read the child's entry, write the parent's; the read of the parent entry is not counted); ncu at the root level: DRAM traffic 8.7 GB = read child + read/write parent, 3.2 TB/s (profiles/r1_ncu_prof_extend_add.txt)"},
                 "solve_roofline": {"bound": "hbm", "kernels": "solve_fwd_step_kernel / solve_bwd_step_kernel (two sweeps over the stored factor)", "achieved": 16.0 * S.stored_nnzL * (1 + refine[0]) / solve_s / 1e9,
                                    "peak": hbm_peak, "unit": "GB/s", "frac": 16.0 * S.stored_nnzL * (1 + refine[0]) / solve_s / 1e9 / hbm_peak,
                                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if hbm_measured else "fallback 6650 GB/s (B200_PROFILING.md)",
                                    "note": "algorithmic bytes = 2 sweeps x 8 B x stored entries of L (panels incl. relaxation zeros); latency-bound on the 256-column steps of the top supernodes, see profiles/r1_solve_launch_summary.txt"}}
    # release the device-resident copy before the plugin path allocates its own
    L.b200_dev_free(ctx, d_vals); L.b200_dev_free(ctx, d_b); L.b200_dev_free(ctx, d_x)
    s.close()

    if world > 1 and not args.no_e2e:
        # ---- e2e at N > 1: the C-ABI shim with HOST buffers on every rank (H2D of the values / b and D2H of x inside the timed calls);
        # the plugin row itself is single-process (SOLVER_SINGLE_THREADED_ONLY), see DESIGN.md 7
        s2 = m.Solver(device=local); s2.join(rank, world, dist=dist)
        s2.analyze(n, cp, ri, m.KIND_CHOLESKY, perm=perm, threads=host_threads)
        ftimes, etimes = [], []
        for i in range(2 + 3):
            dist.barrier(); t = time.perf_counter(); rc = s2.factor(v); tf = time.perf_counter() - t
            t = time.perf_counter(); xe = s2.solve(bvec); te = time.perf_counter() - t
            if rc != 0:
                raise RuntimeError("factor failed: " + s2.err())
            if i >= 2:
                ftimes.append(tf); etimes.append(te)
        import torch
        tt = torch.tensor([float(np.mean(ftimes)), float(np.mean(etimes))], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        berr_e2e = float(np.abs(A @ xe - bvec).max() / (12.0 * np.abs(xe).max() + np.abs(bvec).max()))
        s2.close()
        e2e = {"value": F / tt[0].item() / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(world * (v.nbytes + bvec.nbytes)), "d2h_bytes_per_step": int(world * bvec.nbytes),
               "factorize_s": tt[0].item(), "evaluate_s": tt[1].item(), "backward_error": berr_e2e,
               "path": "b200_factor_host / b200_solve_host (C-ABI shim, pageable host buffers) on every rank, max over ranks"}
    if world == 1 and rank == 0 and not args.no_e2e:
        # ---- e2e through the plugin table with host buffers (what `meagre-crowd -i A.mm -t -s b200` times)
        idx = L.lookup_solver_by_shortname(b"b200")
        A_m = m.make_matrix_csc(n, cp, ri, v, sym=m.SM_SYMMETRIC, location=m.UPPER_TRIANGULAR)
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
        berr_e2e = float(np.abs(A @ xe - bvec).max() / (12.0 * np.abs(xe).max() + np.abs(bvec).max()))
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
                "config": {"workload": f"poisson3d_7pt_{args.grid}^3_spd_cholesky", "step": "factorize + evaluate(p=1, automatic refinement), values and b resident in HBM",
                           "ordering": "nested dissection (own)", "l2": "inputs larger than L2 (factor touches >18 GB per step)",
                           "multi_gpu": f"one factorization over {world} GPUs: subtrees -> ranks, top fronts 1-D block-cyclic, NCCL panels/contribution blocks" if world > 1 else "single GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
