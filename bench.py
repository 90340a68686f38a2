#!/usr/bin/env python
"""Benchmark of the PINN training hot path: residual+gradient collocation points/sec.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU baseline arm, same metric/config)

A step = one full loss + flat-weight-gradient evaluation (jet forward, PDE residual, data misfit,
reverse sweep, cross-GPU reduction of [grad | sums], loss finalisation) over the whole synthetic
collocation set of BASELINE.json configs[4]: 16,777,216 points, [4]+[256]x8+[4] tanh MLP,
Navier_Stokes residual on (t,x,y) with z not differentiated, MSE on the 4 outputs.  The point set
is fixed (strong scaling): with N GPUs every rank owns a contiguous 1/N shard.

The HEADLINE line is the mode that meets north_star's FP32 tolerances (1e-5 loss, 1e-4 gradient)
-- `--precision tf32x3`, the split-operand tensor-core mode -- so that it compares like for like
with the reference's FP32 arithmetic.  The other two modes (plain TF32 operands at a stated looser
bound; the FP32-FMA kernel) are measured in the same run as complete blocks under `modes`.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[4] / SURVEY.md 8(d) primary
    "synthetic16M_256x8_nswe": dict(
        layers=[4] + [256] * 8 + [4], kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2},
        fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3], n=16 * (1 << 20)),
    # SURVEY.md 8(d) secondary (train_newmethod form)
    "synthetic16M_256x8_cont": dict(
        layers=[2] + [256] * 8 + [3], kind="continuity_only", dirs={"x": 0, "y": 1},
        fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1], n=16 * (1 << 20)),
}
CHUNK = 1 << 21   # points per deterministic generation chunk (seed = 1234 + chunk index)
HEADLINE_PRECISION = "tf32x3"

# the reference's own config shapes: name -> (layers, residual, dirs, fields, target cols, N)
_XY = {"x": 0, "y": 1}
_TXY = {"t": 0, "x": 1, "y": 2}
REAL_SHAPES = {
    "config_CMB_h.json": ([2] + [20] * 100 + [3], "continuity_only", _XY, {"U": 0, "V": 1, "h": 2}, [0, 1], 12514),
    "config_CMB.json": ([2] + [10] * 10 + [6], "physics_equation", _XY,
                        {"h": 0, "U": 1, "V": 2, "eta_mean": 3, "Hrms": 4, "k": 5}, [0, 1, 2, 3, 4, 5], 243),
    "config.json": ([5] + [20] * 100 + [4], "Navier_Stokes", _TXY, {"h": 0, "z": 1, "u": 2, "v": 3},
                    [0, 1, 2, 3], 9600),
    "config_txyz.json": ([4] + [20] * 20 + [4], "Navier_Stokes", _TXY, {"h": 0, "z": 1, "u": 2, "v": 3},
                         [0, 1, 2, 3], 9600),
    # the historical physics_functions.Boussinesq residual (third-order input derivatives; bytecode only in the reference)
    "physics_functions.Boussinesq": ([3] + [20] * 20 + [4], "Boussinesq", _TXY, {"h": 0, "z": 1, "u": 2, "v": 3},
                                     [0, 1, 2, 3], 9600),
}

MODE_INFO = {
    "tf32x3": dict(dtype="tf32x3", kernel="pinn::jet_tc_kernel<BWD, X3=true>", bound="tensor",
                   tolerance={"loss_rel": 1e-5, "grad_rel_l2": 1e-4,
                              "mode": "3xTF32 split operands (hi+lo), fp32 accumulate, tanhf: north_star FP32 bound "
                                      "(tests/test_gpu_tc3.py)"}),
    "tf32": dict(dtype="tf32", kernel="pinn::jet_tc_kernel<BWD, X3=false>", bound="tensor",
                 tolerance={"loss_rel": 5e-3, "grad_rel_l2": 5e-3,
                            "mode": "tf32 operands, fp32 accumulate, tanh.approx (stated looser bound; "
                                    "tests/test_gpu_tc.py)"}),
    "fp32": dict(dtype="f32", kernel="pinn::jet_kernel", bound="fp32_fma",
                 tolerance={"loss_rel": 1e-5, "grad_rel_l2": 1e-4, "mode": "fp32 FMA (north_star FP32 bound)"}),
}


def flops_per_point(w):
    """F = 6 (1+k) sum_l in_l*out_l (jet forward 2(1+k)S, reverse 4(1+k)S)."""
    L = w["layers"]
    return 6 * (1 + len(w["dirs"])) * sum(L[i] * L[i + 1] for i in range(len(L) - 1))


def make_shard(w, lo, hi, pin):
    """Synthetic points [lo,hi): inputs U(-1,1), targets N(0,0.05^2); independent of world size."""
    d, nt = w["layers"][0], len(w["target_cols"])
    X = torch.empty(hi - lo, d, dtype=torch.float32, pin_memory=pin)
    T = torch.empty(hi - lo, nt, dtype=torch.float32, pin_memory=pin)
    c0, c1 = lo // CHUNK, (hi + CHUNK - 1) // CHUNK
    for c in range(c0, c1):
        g = torch.Generator().manual_seed(1234 + c)
        xs = torch.rand(CHUNK, d, generator=g) * 2 - 1
        ts = 0.05 * torch.randn(CHUNK, nt, generator=g)
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        X[a - lo:b - lo] = xs[a - c * CHUNK:b - c * CHUNK]
        T[a - lo:b - lo] = ts[a - c * CHUNK:b - c * CHUNK]
    return X, T


def init_params(w):
    """weights of DNN(layers, 0.0, 'xavier') under torch.manual_seed(1234)."""
    from pinn_depthestimation_b200.dnn import DNN
    torch.manual_seed(1234)
    m = DNN(w["layers"], 0.0, "xavier")
    return m.flat_params().clone()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": reasons}


def _oracle_spec(w):
    return dict(layers=w["layers"], activation="tanh", kind=w["kind"], dirs=w["dirs"],
                fields=w["fields"], target_cols=w["target_cols"])


def cpu_port_throughput(w, n_sample, steps, warmup, threads):
    """The reference's CPU algorithm (torch autograd with create_graph, oracle/autograd_port.py)
    on a bounded sample of the same workload; returns (points/s, seconds per step)."""
    from oracle import autograd_port as ap
    torch.set_num_threads(threads)
    spec = _oracle_spec(w)
    X, T = make_shard(w, 0, n_sample, pin=False)
    flat = init_params(w)
    for _ in range(warmup):
        ap.loss_and_grad(spec, flat, X, T)
    t0 = time.perf_counter()
    for _ in range(steps):
        ap.loss_and_grad(spec, flat, X, T)
    dt = (time.perf_counter() - t0) / steps
    return n_sample / dt, dt


def gpu_eager_baseline(w, dev, n_try):
    """The reference's algorithm (eager torch autograd, FP32, TF32 off) on the SAME B200: the in-box GPU
    baseline BASELINE.md section 4 promises.  Largest sample that fits (halved on out-of-memory)."""
    from oracle import autograd_port as ap
    spec = _oracle_spec(w)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    flat = init_params(w).to(dev)
    n = n_try
    try:
        while n >= 1024:
            try:
                Xc, Tc = make_shard(w, 0, n, pin=False)
                X, T = Xc.to(dev), Tc.to(dev)
                ap.loss_and_grad(spec, flat, X, T)       # warm-up (cuBLAS handles, allocator)
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for _ in range(reps):
                    r = ap.loss_and_grad(spec, flat, X, T)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                peak = torch.cuda.max_memory_allocated(dev)
                del X, T, r
                torch.cuda.empty_cache()
                return {"value": n / (ms * 1e-3), "unit": "points/s", "ms_per_eval": ms, "sample_points": n,
                        "dtype": "f32 (allow_tf32=False)", "kind": "port",
                        "what": "oracle/autograd_port.py (eager torch autograd with create_graph, the reference's "
                                "algorithm) on cuda:0; throughput is flat in N beyond ~1e4 points",
                        "peak_mem_gb": peak / 1e9}
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()
                n //= 2
        return {"unavailable": "out of memory down to 1024 points"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def cpu_port_lbfgs(w, n_sample, iters, threads):
    """torch.optim.LBFGS (ctor as train_newmethod.py:108-117) on the reference algorithm, CPU sample."""
    from oracle import autograd_port as ap
    torch.set_num_threads(threads)
    spec = _oracle_spec(w)
    X, T = make_shard(w, 0, n_sample, pin=False)
    p = torch.nn.Parameter(init_params(w))
    opt = torch.optim.LBFGS([p], lr=1, max_iter=iters, max_eval=iters * 5 // 4 + 1, history_size=100,
                            tolerance_grad=1e-5, tolerance_change=1e-7, line_search_fn="strong_wolfe")

    def closure():
        opt.zero_grad()
        r = ap.loss_and_grad(spec, p.detach(), X, T)
        p.grad = r["grad"]
        return r["loss"]
    t0 = time.perf_counter()
    opt.step(closure)
    dt = time.perf_counter() - t0
    st = opt.state[p]
    return {"iterations_per_s": st["n_iter"] / dt, "evaluations_per_s": st["func_evals"] / dt,
            "n_iter": st["n_iter"], "func_evals": st["func_evals"], "ms": dt * 1e3,
            "n_points": n_sample, "history_size": 100, "line_search_fn": "strong_wolfe"}


def run_reference(args, w, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.cpu_points
    v, dt = cpu_port_throughput(w, n_sample, args.steps, args.warmup, threads)
    lb = cpu_port_lbfgs(w, min(n_sample, 16384), args.lbfgs_iters, threads) if args.lbfgs_iters > 0 else None
    line = {
        "impl": "reference", "metric": "residual+grad collocation points/sec", "value": v,
        "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "layers": w["layers"], "residual": w["kind"],
                   "n_points": w["n"], "sample_points_per_step": n_sample},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} of {w['n']} points per step; torch-autograd "
                                   "restatement of dnn.py+physics.py+loss.backward() "
                                   "(oracle/autograd_port.py); /root/reference cannot travel"},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "lbfgs": lb,
    }
    print(json.dumps(line), flush=True)


def measured_traffic(kernel_key):
    """DRAM bytes per point of the dominant kernel from the committed `ncu --set full` capture of THIS build
    (profiles/traffic.json, written by tools/ncu_traffic.py from the .csv export); None when the capture
    belongs to another build of the library, so a stale constant can never be reported."""
    from pinn_depthestimation_b200 import _cabi
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, "no capture committed"
    rec = json.load(open(path)).get(kernel_key)
    ver = _cabi.lib().pinn_version().decode()
    if not rec:
        return None, "no capture for this kernel"
    if rec.get("lib_version") != ver:
        return None, f"capture is of {rec.get('lib_version')!r}, library is {ver!r}"
    return rec, rec.get("source")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synthetic16M_256x8_nswe", choices=list(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override total point count (dev only)")
    ap.add_argument("--precision", default=HEADLINE_PRECISION, choices=list(MODE_INFO))
    ap.add_argument("--other-modes", default="tf32,fp32",
                    help="comma list of further precision modes measured as complete blocks ('' = none)")
    ap.add_argument("--other-steps", type=int, default=3, help="cap on timed steps of the non-headline modes")
    ap.add_argument("--cpu-points", type=int, default=65536)
    ap.add_argument("--eager-points", type=int, default=65536)
    ap.add_argument("--lbfgs-iters", type=int, default=6,
                    help="max_iter of the L-BFGS side measurement (BASELINE metric ii); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--real-shapes", type=int, default=1,
                    help="also time one evaluation of the reference's own config shapes (latency rows)")
    ap.add_argument("--scaling-rows", type=int, default=1,
                    help="n_gpus>1: also time 2^20- and 2^22-point sets (strong scaling where the all-reduce shows)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps
    name = args.workload
    w = dict(WORKLOADS[name])
    if args.points:
        w["n"] = args.points
    if args.impl == "reference":
        return run_reference(args, w, name)

    import ctypes as C
    import torch.distributed as dist
    from pinn_depthestimation_b200 import PassSpec, _cabi
    from pinn_depthestimation_b200.fused import JetLoss

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _cabi.lib()   # fail loudly if the extension is not built

    n_total = w["n"]
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    Xh, Th = make_shard(w, lo, hi, pin=True)
    X, T = Xh.to(dev), Th.to(dev)
    params_h = init_params(w).pin_memory()
    params = params_h.to(dev)
    P = params.numel()
    F = flops_per_point(w)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    grad_h = torch.empty(P, dtype=torch.float32).pin_memory()
    parts_h = torch.empty(4, dtype=torch.float32).pin_memory()

    # ---- roofline denominators, measured in this run -------------------------------------------------
    def fma_peak():
        out = torch.zeros(4, device=dev)
        fl = C.c_double(0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        best = 0.0
        for _ in range(3):
            e0.record()
            _cabi.check(lib.pinn_fma_probe(_cabi.ptr(out), 4096, 148 * 16, C.byref(fl), st))
            e1.record()
            torch.cuda.synchronize()
            best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best, "measured in this run (pinn_fma_probe, FFMA-bound kernel)"

    def tf32_peak():
        bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        old_flag = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ga = torch.randn(8192, 8192, device=dev)
        gb = torch.randn(8192, 8192, device=dev)
        torch.matmul(ga, gb)
        meas = 0.0
        for _ in range(5):
            e0.record()
            torch.matmul(ga, gb)
            e1.record()
            torch.cuda.synchronize()
            meas = max(meas, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = old_flag
        del ga, gb
        return max(meas, bf16 / 2.0), ("max(cuBLAS TF32 GEMM 8192^3 measured in this run = %.1f TFLOP/s, "
                                       "MEASURED_PEAKS.json bf16_tflops_sustained / 2 = %.1f)" % (meas, bf16 / 2.0))

    peak_cache = {}

    def peak_for(bound):
        if bound not in peak_cache:
            peak_cache[bound] = fma_peak() if bound == "fp32_fma" else tf32_peak()
        return peak_cache[bound]

    # ---- one complete measurement of one precision mode ----------------------------------------------
    def measure_mode(precision, steps, warmup, sample_clocks):
        info = MODE_INFO[precision]
        spec = PassSpec(layers=w["layers"], kind=w["kind"], dirs=w["dirs"], fields=w["fields"],
                        target_cols=w["target_cols"], precision=precision)
        jl = JetLoss(spec, X, T, group=group)
        grad = torch.empty_like(params)
        for _ in range(warmup):
            jl.loss_and_grad(params, grad)
        barrier()
        clocks = ClockSampler(local)
        if rank == 0 and sample_clocks:
            clocks.start()
        k0, k1 = [ev() for _ in range(steps)], [ev() for _ in range(steps)]
        barrier()
        e0.record()
        for i in range(steps):
            # the jet kernel alone (+ its 2 memsets and the weight-pack kernel, < 0.01 % of it)
            k0[i].record()
            jl._launch(params, grad, True)
            k1[i].record()
            if world > 1:
                jl._allreduce(grad)
            jl._finalize()
        e1.record()
        barrier()
        clk = clocks.stop() if (rank == 0 and sample_clocks) else None
        step_ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        kern_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in zip(k0, k1)])))
        parts = jl.parts.cpu().numpy()

        # e2e: host buffers in, loss + gradient out, copies inside the timed region
        def e2e_step():
            params.copy_(params_h, non_blocking=True)
            X.copy_(Xh, non_blocking=True)
            T.copy_(Th, non_blocking=True)
            p = jl.loss_and_grad(params, grad)
            grad_h.copy_(grad, non_blocking=True)
            parts_h.copy_(p, non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller reads loss/grad every step

        e2e_step()
        barrier()
        e0.record()
        for _ in range(steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1) / steps)

        peak, peak_src = peak_for(info["bound"])
        achieved = F * (hi - lo) / (kern_ms * 1e-3) / 1e12
        tr, tr_src = measured_traffic(precision)
        roofline = {"bound": info["bound"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak,
                    "traffic": tr["dram_bytes_per_point"] * (hi - lo) if tr else None,
                    "traffic_source": tr_src, "peak_source": peak_src, "kernel": info["kernel"],
                    "kernel_ms": kern_ms, "flops_per_point": F,
                    "points_per_launch": (hi - lo) // max(1, len(jl._shards(jl.res))),
                    "jet_launches_per_evaluation": len(jl._shards(jl.res)),
                    "hbm_gbs_streaming": (hi - lo) * (w["layers"][0] + len(w["target_cols"])) * 4
                    / (kern_ms * 1e-3) / 1e9}
        if precision == "tf32x3":
            # the split-operand mode executes 11 TF32 MMAs per 3 algorithmic ones (4+4 forward/adjoint, 3 weight gradient)
            roofline["executed_tensor_tflops"] = achieved * 11.0 / 3.0
            roofline["executed_frac"] = achieved * 11.0 / 3.0 / peak
            fp32_peak, _ = peak_for("fp32_fma")
            roofline["vs_fp32_fma_peak"] = achieved / fp32_peak
        block = {"precision": precision, "dtype": info["dtype"], "value": n_total / (step_ms * 1e-3),
                 "unit": "points/s", "steps": steps, "warmup": warmup, "ms_per_step": step_ms,
                 "loss_parts": [float(v) for v in parts[:3]],
                 "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                         "h2d_bytes_per_step": (Xh.numel() + Th.numel() + P) * 4, "d2h_bytes_per_step": (P + 4) * 4},
                 "gpu_launches": jl.launches_per_eval * steps, "roofline": roofline, "clocks": clk,
                 "tolerance": info["tolerance"]}
        return block, jl, grad

    head, jl, grad = measure_mode(args.precision, args.steps, args.warmup, True)
    modes = {}
    for m in [s for s in args.other_modes.split(",") if s and s != args.precision]:
        blk, jl_m, grad_m = measure_mode(m, min(args.steps, args.other_steps), 3, True)
        blk["grad_rel_l2_vs_headline"] = float(((grad_m - grad).norm() / grad.norm()).item())
        blk["loss_rel_vs_headline"] = float(abs(blk["loss_parts"][2] - head["loss_parts"][2]) / abs(head["loss_parts"][2]))
        modes[m] = blk
        del jl_m, grad_m
        torch.cuda.empty_cache()

    # ---- second headline metric: L-BFGS iterations/sec (one step(closure) call, train_newmethod.py:204-209) ----
    def lbfgs_run(jl_, p0, iters):
        from pinn_depthestimation_b200.lbfgs import LBFGS
        pl = torch.nn.Parameter(p0.clone())
        opt = LBFGS([pl], lr=1, max_iter=iters, max_eval=iters * 5 // 4 + 1,
                    history_size=100, tolerance_grad=1e-5, tolerance_change=1e-7,
                    line_search_fn="strong_wolfe")      # ctor as train_newmethod.py:108-117

        class _Closure:
            def flat_loss_and_grad(self, fp, fg):
                return jl_.loss_and_grad(fp, fg)
        barrier()
        e0.record()
        opt.step(_Closure())
        e1.record()
        barrier()
        ms_l = max_over_ranks(e0.elapsed_time(e1))
        st_l = opt.state[pl]
        return {"iterations_per_s": st_l["n_iter"] / (ms_l * 1e-3),
                "evaluations_per_s": st_l["func_evals"] / (ms_l * 1e-3),
                "n_iter": st_l["n_iter"], "func_evals": st_l["func_evals"], "ms": ms_l,
                "final_loss": st_l.get("loss"), "history_size": 100, "line_search_fn": "strong_wolfe"}

    lbfgs_side = None
    if args.lbfgs_iters > 0:
        lbfgs_side = lbfgs_run(jl, params, args.lbfgs_iters)
        lbfgs_side["n_points"] = n_total
        lbfgs_side["precision"] = args.precision

    # ---- L-BFGS direction ("HBM-bound, reported as GB/s"): the two whole-chip passes over a full history ----
    def direction_probe(n_par, m):
        al = lambda x: (x + 255) & ~255
        nb = C.c_size_t(0)
        _cabi.check(lib.pinn_lbfgs_workspace_bytes(n_par, m, C.byref(nb)))
        ws = torch.zeros(nb.value + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        tail = 6 * al(4 * n_par) + 2 * al((m + 1) * n_par * 4)          # the [P] vectors and the (s, y) ring
        ws[off + nb.value - tail: off + nb.value].view(torch.float32).normal_(0.0, 1e-3)
        gvec = torch.randn(n_par, device=dev) * 1e-3
        byts = C.c_double(0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        best = None
        for _ in range(4):
            e0.record()
            _cabi.check(lib.pinn_lbfgs_direction_probe(C.c_void_p(ws.data_ptr() + off), n_par, m, _cabi.ptr(gvec),
                                                       C.byref(byts), st))
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1)
            best = ms_ if best is None else min(best, ms_)
        return {"n_params": n_par, "history_size": m, "ms": best, "algorithmic_bytes": byts.value,
                "gb_per_s": byts.value / (best * 1e-3) / 1e9,
                "hbm_peak_gb_per_s": peaks.get("hbm_gbs"),
                "note": "two streaming passes over the 2 m P history (dots with s, y, g; d = sum delta_j b_j) + the "
                        "coefficient-space two-loop in the last CTA; at P = 41,703 the history is L2-resident"}

    lbfgs_direction = None
    if rank == 0 and args.lbfgs_iters > 0:
        lbfgs_direction = [direction_probe(P, 100), direction_probe(41703, 100)]

    # ---- side measurement: the reference's own config shapes (latency + launch count) ----
    real_shapes, lbfgs_real = None, None
    if rank == 0 and args.real_shapes:
        from pinn_depthestimation_b200.dnn import DNN
        real_shapes = []
        for nm, (rl, rk, rd, rf, rt, rn) in REAL_SHAPES.items():
            rspec = PassSpec(layers=rl, kind=rk, dirs=rd, fields=rf, target_cols=rt, precision="fp32")
            g_ = torch.Generator().manual_seed(1234)
            rx = (torch.rand(rn, rl[0], generator=g_) * 2 - 1).to(dev)
            rtg = (0.05 * torch.randn(rn, len(rt), generator=g_)).to(dev)
            torch.manual_seed(1234)
            rp = DNN(rl, 0.0, "xavier").flat_params().clone().to(dev)
            rg = torch.empty_like(rp)
            rj = JetLoss(rspec, rx, rtg)
            for _ in range(3):
                rj.loss_and_grad(rp, rg)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                rj.loss_and_grad(rp, rg)
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1) / 20
            real_shapes.append({"config": nm, "layers": f"[{rl[0]}]+[{rl[1]}]x{len(rl) - 2}+[{rl[-1]}]",
                                "residual": rk, "n_points": rn, "ms_per_eval": ms_,
                                "points_per_s": rn / (ms_ * 1e-3), "kernel_launches_per_eval": 3,
                                "reference_aten_ops_per_eval": "~2900 (SURVEY.md 2.2)"})
            if nm == "config_CMB_h.json" and args.lbfgs_iters > 0 and world == 1:
                # L-BFGS at the reference's real shape (train_newmethod.py:108-117,204-209): 1 ms evaluations, so the
                # optimiser's own vector work and host round trips are what is measured here
                lbfgs_run(rj, rp, 3)     # warm-up (allocations, first launches)
                lbfgs_real = lbfgs_run(rj, rp, 200)
                lbfgs_real.update({"config": nm, "n_points": rn, "precision": "fp32", "ms_per_eval_alone": ms_})

    # ---- the drop-in path: the body of pinn.loss_func (train_newmethod.py:120-159) + loss.backward() on dropin/{dnn,physics}.py ----
    dropin_row = None
    if rank == 0 and args.real_shapes:
        sys.path.insert(0, os.path.join(ROOT, "dropin"))
        import dnn as dropin_dnn
        import physics as dropin_physics
        rl, rn = [2] + [20] * 100 + [3], 12514
        torch.manual_seed(1234)
        mdl = dropin_dnn.DNN(rl, 0.0, "xavier").to(dev)
        g_ = torch.Generator().manual_seed(1234)
        xs = (torch.rand(rn, 2, generator=g_) * 2 - 1)
        xq = xs[:, 0:1].clone().to(dev).requires_grad_(True)
        yq = xs[:, 1:2].clone().to(dev).requires_grad_(True)
        tq = (0.05 * torch.randn(rn, 2, generator=g_)).to(dev)

        def dropin_step():
            for p_ in mdl.parameters():
                p_.grad = None
            pred = mdl(torch.cat([xq, yq], dim=-1))
            fid = sum(torch.nn.functional.mse_loss(pred[:, i:i + 1], tq[:, i:i + 1]) for i in range(2))
            res = dropin_physics.continuity_only(xq, yq, pred[:, 2:3], pred[:, 0:1], pred[:, 1:2])
            (fid + res).backward()
        for _ in range(3):
            dropin_step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            dropin_step()
        e1.record()
        torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 20
        dropin_row = {"config": "config_CMB_h.json", "n_points": rn, "ms_per_loss_func_and_backward": ms_,
                      "points_per_s": rn / (ms_ * 1e-3),
                      "note": "unmodified loss_func body on the drop-in modules: value forward + fused residual fwd/bwd + "
                              "external-seed fwd/bwd for the MSE part (3 network evaluations); the fused trainer does one"}

    # ---- strong-scaling rows at 2^20 and 2^22 points: where the one all-reduce becomes visible ----
    scaling_rows = None
    if world > 1 and args.scaling_rows:
        scaling_rows = []
        for n_small in (1 << 20, 1 << 22):
            l2, h2 = rank * n_small // world, (rank + 1) * n_small // world
            spec = PassSpec(layers=w["layers"], kind=w["kind"], dirs=w["dirs"], fields=w["fields"],
                            target_cols=w["target_cols"], precision=args.precision)
            js = JetLoss(spec, X[:h2 - l2].contiguous(), T[:h2 - l2].contiguous(), group=group)
            gs = torch.empty_like(params)
            for _ in range(3):
                js.loss_and_grad(params, gs)
            barrier()
            e0.record()
            for _ in range(10):
                js.loss_and_grad(params, gs)
            e1.record()
            barrier()
            ms_s = max_over_ranks(e0.elapsed_time(e1) / 10)
            scaling_rows.append({"n_points": n_small, "ms_per_step": ms_s, "value": n_small / (ms_s * 1e-3),
                                 "unit": "points/s", "precision": args.precision})
            del js, gs

    line = None
    if rank == 0:
        line = {
            "metric": "residual+grad collocation points/sec", "value": head["value"], "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": head["dtype"], "data": "synthetic",
            "config": {"workload": name, "layers": w["layers"], "residual": w["kind"],
                       "n_points": n_total, "points_per_gpu": hi - lo, "parallelism": f"dp{world}",
                       "precision": args.precision,
                       "l2_policy": "inputs (%.0f MB per GPU) larger than L2, not flushed"
                                    % ((Xh.numel() + Th.numel()) * 4 / 1e6)},
            "loss_parts": head["loss_parts"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"],
            "gpu_launches_note": "per step: pack kernel, jet kernel (one per slice of at most fused.SHARD_POINTS points: the "
                                 "FP32 kernel's 16.8M points run as 8 launches so that no FP32 running sum sees more than "
                                 "2M points' tiles), finalize_kernel (+2 memsets, "
                                 "+1 NCCL all-reduce when n_gpus>1)",
            "roofline": head["roofline"], "clocks": head["clocks"],
            "tolerance": head["tolerance"],
            "modes": modes,
            "lbfgs": lbfgs_side,
            "lbfgs_real_shape": lbfgs_real,
            "lbfgs_direction": lbfgs_direction,
            "dropin_real_shape": dropin_row,
            "real_shapes_fp32": real_shapes,
            "strong_scaling_rows": scaling_rows,
        }
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        line["gpu_eager_baseline"] = gpu_eager_baseline(w, dev, args.eager_points)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_port_throughput(w, args.cpu_points, 2, 1, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": "points/s", "cores": threads, "kind": "port",
            "sample": f"{args.cpu_points} of {n_total} points, 2 timed evaluations after 1 warm-up, "
                      f"{dt:.2f} s each; torch-autograd restatement of the reference path"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
