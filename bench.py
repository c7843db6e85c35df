#!/usr/bin/env python
"""bench.py — Arnoldi iterations/s and HBM roofline on BASELINE.json's headline config.

A "step" is one complete 200-iteration CGS2 Arnoldi cycle of the hybrid AB-/BA-GMRES
hot path (operator B*(A*q)+lambda*q, two-pass CGS, normalise) on the 1024x1024
fan-beam phantom problem with an unmatched pixel-driven back-projector (BASELINE.json
configs[3]); `value` = 200*K / device time.  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # CPU-baseline leg: see oracle/cport.py

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, n_views, geometry, maxit)
    "ct1024_fan_180v_pixelB_k200": dict(N=1024, n_views=180, geometry="fan", maxit=200),
    "ct512_fan_180v_pixelB_k200": dict(N=512, n_views=180, geometry="fan", maxit=200),
    "ct256_fan_180v_pixelB_k100": dict(N=256, n_views=180, geometry="fan", maxit=100),
    "ct64_par_180v_pixelB_k80": dict(N=64, n_views=180, geometry="parallel", maxit=80),
    # BASELINE configs[4]: 2048^2, 3600 views x 2896 detectors (19 G + 30 G non-zeros): 8 GPUs only
    "ct2048_par_3600v_pixelB_k200": dict(N=2048, n_views=3600, geometry="parallel", maxit=200),
}
LAMBDA = 1e-2  # run_2D_phantom.m:8
NOISE = 0.01   # BASELINE config 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_workload(hg, ctx, name, rank=0, world=1, dist=None, order="natural"):
    """Matrices are generated on the device.  world > 1: this rank's detector rows (whole views) A_p and
    the matching columns B^p; b is the matching part of the sinogram."""
    w = WORKLOADS[name]
    N, nv, geom = w["N"], w["n_views"], w["geometry"]
    angles = np.arange(nv) * ((360.0 if geom == "fan" else 180.0) / nv)
    p = int(round(math.sqrt(2.0) * N))
    from hybrid_gmres_b200.ct import ct_backprojector, ct_projector, shepp_logan
    m = nv * p
    # rank r owns the views r, r+P, r+2P, ... (whole detector rows of the sinogram): every rank sees the
    # same mix of ray directions, so the gather-bound projector product takes the same time on every rank
    # (contiguous view blocks left ranks waiting ~10 % of the step at 2048^2) and every view carries the
    # same number of non-zeros, so the blocks are nnz balanced
    mine = np.arange(rank, nv, world)
    dA = ct_projector(N, angles[mine], p, geom, ctx=ctx)
    dB = ct_backprojector(N, angles[mine], p, geom, ctx=ctx)
    x_true = shepp_logan(N)
    b_exact = dA.matvec(x_true)
    rng = np.random.default_rng(0)
    e = rng.standard_normal(m)
    nb2 = float(b_exact @ b_exact)
    if world > 1:
        import torch
        t = torch.tensor([nb2], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        nb2 = float(t.item())
    e_mine = e.reshape(nv, p)[mine].ravel()
    b = b_exact + NOISE * math.sqrt(nb2) * e_mine / np.linalg.norm(e)  # run_2D_phantom.m:18-19
    if order.startswith("tile"):
        # n-space in tile x tile pixel blocks (hg_matrix_permute): A(:,q), B(q,:), x_true(q)
        from hybrid_gmres_b200.ct import tile_permutation
        q = tile_permutation(N, int(order[4:]))
        dA2 = dA.permute(None, q, sort=False)
        dA.close()
        ctx.trim()  # the shards of configs[4] are tens of GB: do not keep released copies cached
        dB2 = dB.permute(q, None)
        dB.close()
        ctx.trim()
        dA, dB, x_true = dA2, dB2, np.ascontiguousarray(x_true[q])
    return dA, dB, b, x_true, w["maxit"]


def host_csr(dM):
    import scipy.sparse as sp
    indptr, indices, data = dM.download()
    return sp.csr_matrix((data, indices, indptr), shape=dM.shape)


def cpu_reference_sample(A, B, b, x_true, iters):
    """The reference's BA-RTP loop (hybrid_ba_gmres_rtp.m restated literally) for `iters`
    iterations on the host cores: the OpenMP C restatement (all cores) when it has been built,
    else the NumPy/SciPy one.  Returns (iters/s, seconds, iters, description, threads)."""
    from oracle import cport
    if cport.available():
        t0 = time.perf_counter()
        x, err, res, it = cport.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, iters, LAMBDA)
        dt = time.perf_counter() - t0
        th = cport.num_threads()
        return it / dt, dt, it, f"oracle/c/hg_oracle.c (OpenMP, {th} threads)", th
    import oracle
    t0 = time.perf_counter()
    x, err, res, it = oracle.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, iters, LAMBDA)
    dt = time.perf_counter() - t0
    return it / dt, dt, it, "oracle/solvers.py (SciPy CSR mat-vec single-threaded, BLAS threaded)", 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ct1024_fan_180v_pixelB_k200", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-iters", type=int, default=20, help="iterations of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nspace-order", default="tile4", choices=["natural", "tile4", "tile8"],
                    help="pixel order of the n-space on the device (hg_matrix_permute)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(args.steps, 1), max(args.warmup, 3)  # never fewer than 3 warm-up cycles (reported as run)

    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    import hybrid_gmres_b200 as hg

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # The library launches on the stream it is given; torch's events only see torch's current stream.  The
    # legacy default stream has handle 0 (the library would then create its own non-blocking stream, which
    # events on the default stream do not wait for), so make an explicit stream current and hand it over.
    bench_stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(bench_stream)
    stream = bench_stream.cuda_stream
    assert stream != 0
    ctx = hg.Context(local_rank, stream=stream)
    if args.impl == "reference":
        dA, dB, b, x_true, maxit = build_workload(hg, ctx, args.workload)
    else:
        dA, dB, b, x_true, maxit = build_workload(hg, ctx, args.workload, rank, world, dist, args.nspace_order)
    m, n = dA.shape
    nnzA, nnzB = dA.nnz, dB.nnz
    if dist is not None:
        t = torch.tensor([float(m), float(nnzA), float(nnzB)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        m, nnzA, nnzB = (int(v) for v in t.tolist())
    cores = len(os.sched_getaffinity(0))
    config = {"workload": args.workload, "m": m, "n": n, "nnz_A": nnzA, "nnz_B": nnzB, "maxit": maxit,
              "lambda": LAMBDA, "orth": "cgs2", "B": "pixel-driven (unmatched)",
              "nspace_order": args.nspace_order,
              "spmv_form": {"A": f"{dA.spmv_form}/idx{dA.spmv_index_bits}", "B": f"{dB.spmv_form}/idx{dB.spmv_index_bits}"},
              "parallelism": (f"A row-sharded / B column-sharded x{world}, NCCL reduce-scatter + all-gather + "
                              "3 all-reduce per step") if world > 1 else "single",
              "l2": "inputs (A+B = %.1f GB) exceed the 126 MB L2; no flush needed" % ((nnzA + nnzB) * 12 / 1e9)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        A, B = host_csr(dA), host_csr(dB)
        del dA, dB
        for _ in range(W):
            cpu_reference_sample(A, B, b, x_true, 2)
        t0 = time.perf_counter()
        its = 0
        for _ in range(K):
            _, _, it, desc, threads = cpu_reference_sample(A, B, b, x_true, args.cpu_iters)
            its += it
        dt = time.perf_counter() - t0
        val = its / dt
        line = {"impl": "reference", "metric": "arnoldi_iters_per_s", "value": val, "unit": "iter/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "iter/s", "cores": threads, "host_cores": cores, "kind": "port",
                                 "sample": f"{args.cpu_iters} full hybrid BA-RTP iterations per step of the same workload "
                                           f"(hybrid_ba_gmres_rtp.m restated literally: 3 SpMV + MGS + projected LS per "
                                           f"iteration) by {desc}"},
                "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    sharded = world > 1
    if sharded:
        from hybrid_gmres_b200.distributed import Communicator, ShardedArnoldi
        comm = Communicator(ctx)
        ar = ShardedArnoldi(comm, dA, dB, maxit)
        ar.set_rhs(b)
        peer = comm.transport.startswith("peer")
        config["transport"] = comm.transport
        config["parallelism"] = (
            f"A row-sharded by whole views (rank r: views r, r+{world}, ...) / B column-sharded x{world}; per step: reduce-scatter pulled over NVLink peer memory by "
            "the first CGS2 multi-dot, 3 one-shot all-reduces inside the second-stage reductions, all-gather "
            "pushed by the normalisation kernel (no NCCL call on the step)" if peer else
            f"A row-sharded / B column-sharded x{world}, NCCL reduce-scatter + all-gather + 3 all-reduce per step")
    else:
        ar = hg.Arnoldi(dA, dB, "n", maxit)
        ar.set_rhs(b)
    for _ in range(W):
        ar.reset(LAMBDA)
        ar.steps(maxit)
    ctx.sync()

    def timed_pass(instrument):
        """K steps bracketed by barrier + synchronize; device time from CUDA events on the launch stream.
        instrument=True additionally records an event pair around every kernel launch (hg_ctx_timing_*)."""
        ctx.timing_enable(instrument)
        ctx.timing_reset()
        l0 = ctx.launch_count
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(K):
            ar.reset(LAMBDA)
            ar.steps(maxit)
        ev1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        clk = sampler.stop()
        t_ms = ev0.elapsed_time(ev1)
        nl = ctx.launch_count - l0
        tim = ctx.timing() if instrument else None
        ctx.timing_enable(False)
        if dist is not None:
            t = torch.tensor([t_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, nl, clk, tim

    # `value` comes from the un-instrumented pass; the per-kernel CUDA-event times of the roofline
    # come from a second, instrumented pass over the same K steps (two event records per launch
    # serialise back-to-back kernels a little, so its step time is reported next to the clean one)
    ms, launches, clocks, _ = timed_pass(False)
    ms_instr, _, _, timing = timed_pass(True)
    total_iters = maxit * K  # one sharded job: the same 200-iteration cycle on N GPUs (strong scaling)
    value = total_iters / (ms * 1e-3)
    step_bytes = sum(ar.step_bytes(k) for k in range(1, maxit + 1))
    if sharded:  # whole-job algorithmic bytes = sum over ranks
        t = torch.tensor([step_bytes], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        step_bytes = float(t.item())
        peak_scale = world
    else:
        peak_scale = 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    sp_ms, sp_cnt, sp_bytes = timing["spmv"]
    achieved = (sp_bytes / sp_cnt) / (sp_ms / sp_cnt * 1e-3) / 1e9 if sp_cnt else 0.0
    traffic, traffic_src = None, None
    if args.workload == "ct1024_fan_180v_pixelB_k200":
        try:  # dram__bytes_read+write per launch from the committed `ncu --set full` capture of this kernel
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_spmv_traffic.json")))
            traffic, traffic_src = float(tj["mean"]), tj["source"]
        except Exception:
            pass
    forms = config.get("spmv_form", {})
    knames = {"csr/idx32": "spmv_csr_kernel<32>", "csr/idx16": "spmv_csr16_kernel", "sell32/idx32": "spmv_sell32_kernel<4>",
              "sell32/idx16": "spmv_sell16_kernel<4>", "stream/idx32": "spmv_stream_kernel"}
    kname = " + ".join(f"{knames.get(forms.get(w), 'spmv')} ({w})" for w in ("A", "B"))
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "launches": sp_cnt, "avg_launch_ms": sp_ms / sp_cnt if sp_cnt else None,
                "algorithmic_bytes_per_launch": sp_bytes / sp_cnt if sp_cnt else None,
                "instrumented_ms_per_step": ms_instr / K,
                "step_algorithmic_GBps": step_bytes * K / (ms * 1e-3) / 1e9,
                "step_frac": step_bytes * K / (ms * 1e-3) / 1e9 / (peak * peak_scale),
                "per_class": {k: {"ms": v[0], "launches": v[1],
                                  "GBps": (v[2] / (v[0] * 1e-3) / 1e9) if v[0] > 0 else None}
                              for k, v in timing.items() if v[1]}}
    line = {"metric": "arnoldi_iters_per_s", "value": value, "unit": "iter/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline,
            "gpu_launches": launches, "clocks": clocks}

    # ------------------------------------------------------------------ e2e through the public API
    if sharded and not args.no_e2e:
        # end to end of the sharded public API: every rank uploads its shards (natural pixel order) from
        # pinned host memory; the sharded hybrid BA-GMRES re-orders the n-space on the device, runs 200
        # full hybrid iterations (Arnoldi + projected solve + iterate + both histories) and returns x
        from hybrid_gmres_b200 import distributed as hgd
        nperm = None
        if args.nspace_order.startswith("tile"):
            from hybrid_gmres_b200.ct import tile_permutation
            nperm = tile_permutation(WORKLOADS[args.workload]["N"], int(args.nspace_order[4:]))
            qinv = np.empty_like(nperm)
            qinv[nperm] = np.arange(nperm.shape[0], dtype=nperm.dtype)
            nA, nB = dA.permute(None, qinv), dB.permute(qinv, None)
            ar.close()
            dA.close()
            dB.close()
            dA, dB = nA, nB
            xt = np.empty_like(x_true)
            xt[nperm] = x_true
            x_true = xt
        A, B = host_csr(dA), host_csr(dB)
        pinned = []
        for arr in (A.indptr, A.indices, A.data, B.indptr, B.indices, B.data, b, x_true):
            try:
                hg._lib.check(ctx._lib.hg_host_register(arr.ctypes.data, arr.nbytes))
                pinned.append(arr)
            except Exception:
                pass
        ar.close()
        dA.close()
        dB.close()
        del ar, dA, dB
        hgd.hybrid_ba_gmres_rtp(comm, A, B, b, x_true, 0.0, 3, LAMBDA, nperm=nperm)  # warm-up (buffer cache)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Ke = max(1, min(K, 3))
        its = 0
        for _ in range(Ke):
            x, err, res, it = hgd.hybrid_ba_gmres_rtp(comm, A, B, b, x_true, 0.0, maxit, LAMBDA, nperm=nperm)
            its += it
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = A.indptr.nbytes + A.indices.nbytes + A.data.nbytes + B.indptr.nbytes + B.indices.nbytes + \
            B.data.nbytes + b.nbytes + x_true.nbytes
        tb = torch.tensor([float(h2d)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tb)
        line["e2e"] = {"value": its / float(t.item()), "unit": "iter/s", "h2d_bytes_per_step": int(tb.item()),
                       "d2h_bytes_per_step": int(world * (x.nbytes + maxit * (maxit + 3) // 2 * 8 + maxit * 16)),
                       "steps": Ke, "final_residual": float(res[-1]),
                       "api": "distributed.hybrid_ba_gmres_rtp(comm, A_p, B_p, b_p, x_true, tol, maxit, lambda): per-rank "
                              "upload of the shards from pinned host CSR (natural pixel order) + device re-ordering + "
                              "200 full sharded hybrid iterations + histories per step"}
        for arr in pinned:
            ctx._lib.hg_host_unregister(arr.ctypes.data)
    elif rank == 0 and not args.no_e2e:
        nperm = None
        if args.nspace_order.startswith("tile"):
            # the caller's host data is in the natural (column-major pixel) order; the library applies the
            # tile order on the device inside the timed call (nperm=) and returns x in the caller's order
            from hybrid_gmres_b200.ct import tile_permutation
            nperm = tile_permutation(WORKLOADS[args.workload]["N"], int(args.nspace_order[4:]))
            qinv = np.empty_like(nperm)
            qinv[nperm] = np.arange(nperm.shape[0], dtype=nperm.dtype)
            nA, nB = dA.permute(None, qinv), dB.permute(qinv, None)
            ar.close()
            dA.close()
            dB.close()
            dA, dB = nA, nB
            xt = np.empty_like(x_true)
            xt[nperm] = x_true
            x_true = xt
        A, B = host_csr(dA), host_csr(dB)
        pinned = []
        for arr in (A.indptr, A.indices, A.data, B.indptr, B.indices, B.data, b, x_true):
            try:
                hg._lib.check(ctx._lib.hg_host_register(arr.ctypes.data, arr.nbytes))
                pinned.append(arr)
            except Exception as exc:  # pageable upload still works, only slower
                print(f"[bench] hg_host_register failed for a {arr.nbytes}-byte array: {exc}", file=sys.stderr)
        ar.close()
        del ar
        Ke = max(1, min(K, 3))
        hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, min(maxit, 3), LAMBDA, ctx=ctx, nperm=nperm)  # warm-up (allocators)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        its = 0
        step_s = []
        for _ in range(Ke):
            ts = time.perf_counter()
            x, err, res, it = hg.hybrid_ba_gmres_rtp(A, B, b, x_true, 0.0, maxit, LAMBDA, ctx=ctx, nperm=nperm)
            its += it
            step_s.append(round(time.perf_counter() - ts, 4))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = A.indptr.nbytes + A.indices.nbytes + A.data.nbytes + B.indptr.nbytes + B.indices.nbytes + \
            B.data.nbytes + b.nbytes + x_true.nbytes + maxit * (maxit + 1) // 2 * 8
        d2h = x.nbytes + maxit * (maxit + 3) // 2 * 8 + maxit * 16
        line["e2e"] = {"value": its / dt, "unit": "iter/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "steps": Ke, "step_s": step_s,
                       "final_residual": float(res[-1]),
                       "api": "hybrid_ba_gmres_rtp(A,B,b,x_true,tol,maxit,lambda) with pinned host CSR A, B in the "
                              "caller's natural pixel order (upload + device re-ordering + 200 full hybrid "
                              "iterations + histories per step)"}
        t0 = time.perf_counter()
        x, err, res, it = hg.hybrid_ab_gmres_rtp(A, B, b, x_true, 0.0, maxit, LAMBDA, ctx=ctx, nperm=nperm)
        torch.cuda.synchronize()
        line["e2e"]["hybrid_ab_iters_per_s"] = it / (time.perf_counter() - t0)
        if not args.no_cpu:
            v, dt, it, desc, threads = cpu_reference_sample(A, B, b, x_true, args.cpu_iters)
            line["cpu_baseline"] = {"value": v, "unit": "iter/s", "cores": threads, "host_cores": cores, "kind": "port",
                                    "sample": f"{it} full hybrid BA-RTP iterations of the same workload ({dt:.1f} s) by "
                                              f"{desc} — compare with e2e (full hybrid iterations), not with value"}
        for arr in pinned:
            ctx._lib.hg_host_unregister(arr.ctypes.data)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
