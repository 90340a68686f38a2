#!/usr/bin/env python
"""Benchmark of the PINN training hot path: residual+gradient collocation points/sec.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU baseline arm, same metric/config)

A step = one full loss + flat-weight-gradient evaluation (jet forward, PDE residual, data misfit,
reverse sweep, cross-GPU reduction of [grad | sums], loss finalisation) over the whole synthetic
collocation set of BASELINE.json configs[4]: 16,777,216 points, [4]+[256]x8+[4] tanh MLP,
Navier_Stokes residual on (t,x,y) with z not differentiated, MSE on the 4 outputs.  The point set
is fixed (strong scaling): with N GPUs every rank owns a contiguous 1/N shard.

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


def cpu_port_throughput(w, n_sample, steps, warmup, threads):
    """The reference's CPU algorithm (torch autograd with create_graph, oracle/autograd_port.py)
    on a bounded sample of the same workload; returns (points/s, seconds per step)."""
    from oracle import autograd_port as ap
    from oracle import jet_oracle as jo
    torch.set_num_threads(threads)
    spec = dict(layers=w["layers"], activation="tanh", kind=w["kind"], dirs=w["dirs"],
                fields=w["fields"], target_cols=w["target_cols"])
    assert spec["kind"] in (jo.NSWE, jo.CONT_ONLY)
    X, T = make_shard(w, 0, n_sample, pin=False)
    flat = init_params(w)
    for _ in range(warmup):
        ap.loss_and_grad(spec, flat, X, T)
    t0 = time.perf_counter()
    for _ in range(steps):
        ap.loss_and_grad(spec, flat, X, T)
    dt = (time.perf_counter() - t0) / steps
    return n_sample / dt, dt


def cpu_port_lbfgs(w, n_sample, iters, threads):
    """torch.optim.LBFGS (ctor as train_newmethod.py:108-117) on the reference algorithm, CPU sample."""
    from oracle import autograd_port as ap
    torch.set_num_threads(threads)
    spec = dict(layers=w["layers"], activation="tanh", kind=w["kind"], dirs=w["dirs"],
                fields=w["fields"], target_cols=w["target_cols"])
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
    lb = cpu_port_lbfgs(w, n_sample, args.lbfgs_iters, threads) if args.lbfgs_iters > 0 else None
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synthetic16M_256x8_nswe", choices=list(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override total point count (dev only)")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--fp32-steps", type=int, default=2, help="timed steps of the FP32 parity-mode side measurement (0 = skip)")
    ap.add_argument("--cpu-points", type=int, default=16384)
    ap.add_argument("--lbfgs-iters", type=int, default=6,
                    help="max_iter of the L-BFGS side measurement (BASELINE metric ii); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--real-shapes", type=int, default=1,
                    help="also time one evaluation of the reference's own config shapes (latency rows)")
    args = ap.parse_args()
    name = args.workload
    w = dict(WORKLOADS[name])
    if args.points:
        w["n"] = args.points
    if args.impl == "reference":
        return run_reference(args, w, name)

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
    _cabi.lib()   # fail loudly if the extension is not built

    n_total = w["n"]
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    Xh, Th = make_shard(w, lo, hi, pin=True)
    X, T = Xh.to(dev), Th.to(dev)
    params_h = init_params(w).pin_memory()
    params = params_h.to(dev)
    grad = torch.empty_like(params)
    spec = PassSpec(layers=w["layers"], kind=w["kind"], dirs=w["dirs"], fields=w["fields"],
                    target_cols=w["target_cols"], precision=args.precision)
    jl = JetLoss(spec, X, T, group=group)
    P = params.numel()

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

    # ---- value: inputs resident in HBM -------------------------------------------------------
    for _ in range(args.warmup):
        jl.loss_and_grad(params, grad)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    k0, k1 = [ev() for _ in range(args.steps)], [ev() for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    for i in range(args.steps):
        # the jet kernel alone (+ its 2 memsets and the weight-pack kernel, < 0.01 % of it)
        k0[i].record()
        jl._launch(params, grad, True)
        k1[i].record()
        if world > 1:
            jl._allreduce(grad)
        jl._finalize()
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    step_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    kern_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in zip(k0, k1)])))
    parts = jl.parts.cpu().numpy()
    value = n_total / (step_ms * 1e-3)

    # ---- e2e: host buffers in, loss + gradient out, copies inside the timed region -----------
    grad_h = torch.empty(P, dtype=torch.float32).pin_memory()
    parts_h = torch.empty(4, dtype=torch.float32).pin_memory()

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
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    h2d = (Xh.numel() + Th.numel() + P) * 4
    d2h = (P + 4) * 4

    # ---- second headline metric: L-BFGS iterations/sec (one step(closure) call, train_newmethod.py:204-209) ----
    lbfgs_side = None
    if args.lbfgs_iters > 0:
        from pinn_depthestimation_b200.lbfgs import LBFGS
        pl = torch.nn.Parameter(params.clone())
        opt = LBFGS([pl], lr=1, max_iter=args.lbfgs_iters, max_eval=args.lbfgs_iters * 5 // 4 + 1,
                    history_size=100, tolerance_grad=1e-5, tolerance_change=1e-7,
                    line_search_fn="strong_wolfe")      # ctor as train_newmethod.py:108-117

        class _Closure:
            def flat_loss_and_grad(self, fp, fg):
                return jl.loss_and_grad(fp, fg)
        barrier()
        e0.record()
        opt.step(_Closure())
        e1.record()
        barrier()
        ms_l = max_over_ranks(e0.elapsed_time(e1))
        st_l = opt.state[pl]
        lbfgs_side = {"iterations_per_s": st_l["n_iter"] / (ms_l * 1e-3),
                      "evaluations_per_s": st_l["func_evals"] / (ms_l * 1e-3),
                      "n_iter": st_l["n_iter"], "func_evals": st_l["func_evals"], "ms": ms_l,
                      "final_loss": st_l.get("loss"), "history_size": 100,
                      "line_search_fn": "strong_wolfe", "n_points": n_total}
        del opt, pl

    # ---- side measurement: the reference's own config shapes (latency + launch count) ----
    real_shapes = None
    if rank == 0 and args.real_shapes:
        real_shapes = []
        for nm, (rl, rk, rd, rf, rt, rn) in REAL_SHAPES.items():
            rspec = PassSpec(layers=rl, kind=rk, dirs=rd, fields=rf, target_cols=rt, precision="fp32")
            g_ = torch.Generator().manual_seed(1234)
            rx = (torch.rand(rn, rl[0], generator=g_) * 2 - 1).to(dev)
            rtg = (0.05 * torch.randn(rn, len(rt), generator=g_)).to(dev)
            torch.manual_seed(1234)
            from pinn_depthestimation_b200.dnn import DNN
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

    # ---- side measurement: the FP32 parity mode on the same workload ---------------------------
    fp32_side = None
    if args.precision != "fp32" and args.fp32_steps > 0:
        spec32 = PassSpec(layers=w["layers"], kind=w["kind"], dirs=w["dirs"], fields=w["fields"],
                          target_cols=w["target_cols"], precision="fp32")
        jl32 = JetLoss(spec32, X, T, group=group)
        g32 = torch.empty_like(params)
        jl32.loss_and_grad(params, g32)
        barrier()
        e0.record()
        for _ in range(args.fp32_steps):
            jl32.loss_and_grad(params, g32)
        e1.record()
        barrier()
        ms32 = max_over_ranks(e0.elapsed_time(e1) / args.fp32_steps)
        p32 = jl32.parts.cpu().numpy()
        fp32_side = {"value": n_total / (ms32 * 1e-3), "unit": "points/s", "ms_per_step": ms32,
                     "steps": args.fp32_steps, "warmup": 1, "loss_parts": [float(v) for v in p32[:3]],
                     "tflops": flops_per_point(w) * n_total / (ms32 * 1e-3) / 1e12,
                     "grad_rel_l2_vs_headline": float(((grad - g32).norm() / g32.norm()).item()),
                     "loss_rel_vs_headline": float(abs(parts[2] - p32[2]) / abs(p32[2]))}
        del jl32, g32

    # ---- roofline denominators -----------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    F = flops_per_point(w)
    achieved = F * (hi - lo) / (kern_ms * 1e-3) / 1e12
    if args.precision == "fp32":
        # FP32-FMA peak is not in MEASURED_PEAKS.json: measure it here with the library's probe
        out = torch.zeros(4, device=dev)
        import ctypes as C
        fl = C.c_double(0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib = _cabi.lib()
        best = 0.0
        for _ in range(3):
            e0.record()
            _cabi.check(lib.pinn_fma_probe(_cabi.ptr(out), 4096, 148 * 16, C.byref(fl), st))
            e1.record()
            torch.cuda.synchronize()
            best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        peak, bound, peak_src = best, "fp32_fma", "measured in this run (pinn_fma_probe, FFMA-bound kernel)"
    else:
        # TF32 tensor peak: MEASURED_PEAKS.json holds only bf16, so measure a cuBLAS TF32 GEMM here
        # (8192^3, best of 5) and keep the bf16-derived figure beside it
        bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        old_flag = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ga = torch.randn(8192, 8192, device=dev)
        gb = torch.randn(8192, 8192, device=dev)
        torch.matmul(ga, gb)
        tf32_meas = 0.0
        for _ in range(5):
            e0.record()
            torch.matmul(ga, gb)
            e1.record()
            torch.cuda.synchronize()
            tf32_meas = max(tf32_meas, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = old_flag
        del ga, gb
        peak, bound = max(tf32_meas, bf16 / 2.0), "tensor"
        peak_src = ("max(cuBLAS TF32 GEMM 8192^3 measured in this run = %.1f TFLOP/s, "
                    "MEASURED_PEAKS.json bf16_tflops_sustained / 2 = %.1f)" % (tf32_meas, bf16 / 2.0))
    # DRAM bytes per point from the committed `ncu --set full` captures of a 1,048,576-point launch
    # (profiles/r1_fp32_v7_*, r1_tf32_v5_*_ncu_full.csv: dram__bytes_read.sum + dram__bytes_write.sum), scaled to this launch
    dram_per_point = {"fp32": (0.900642048e9 + 28.083769e9) / 1048576,   # profiles/r1_fp32_v7_*
                      "tf32": (23.591678e9 + 61.434800e9) / 1048576}[args.precision]   # profiles/r1_tf32_v5_*
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": dram_per_point * (hi - lo),
                "traffic_note": "bytes; ncu capture of a 1,048,576-point launch scaled by points",
                "peak_source": peak_src,
                "kernel": "pinn::jet_kernel" if args.precision == "fp32" else "pinn::jet_tc_kernel", "kernel_ms": kern_ms,
                "flops_per_point": F, "points_per_launch": hi - lo,
                "hbm_gbs_streaming": (hi - lo) * (w["layers"][0] + len(w["target_cols"])) * 4
                / (kern_ms * 1e-3) / 1e9}

    line = None
    if rank == 0:
        line = {
            "metric": "residual+grad collocation points/sec", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": name, "layers": w["layers"], "residual": w["kind"],
                       "n_points": n_total, "points_per_gpu": hi - lo, "parallelism": f"dp{world}",
                       "l2_policy": "inputs (%.0f MB per GPU) larger than L2, not flushed"
                                    % ((Xh.numel() + Th.numel()) * 4 / 1e6)},
            "loss_parts": [float(v) for v in parts[:3]],
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 3 * args.steps,
            "gpu_launches_note": "per step: pack_kernel, jet_kernel, finalize_kernel (+2 memsets, "
                                 "+1 NCCL all-reduce when n_gpus>1)",
            "roofline": roofline, "clocks": clk,
            "tolerance": ({"loss_rel": 1e-5, "grad_rel_l2": 1e-4, "mode": "fp32 (north_star FP32 bound)"}
                          if args.precision == "fp32" else
                          {"loss_rel": 3e-3, "grad_rel_l2": 5e-3,
                           "mode": "tf32 operands, fp32 accumulate (stated looser bound; tests/test_gpu_tc.py)"}),
            "fp32_parity_mode": fp32_side,
            "lbfgs": lbfgs_side,
            "real_shapes_fp32": real_shapes,
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_port_throughput(w, args.cpu_points, 3, 1, threads)
        line["cpu_baseline"] = {
            "value": v, "unit": "points/s", "cores": threads, "kind": "port",
            "sample": f"{args.cpu_points} of {n_total} points, 3 timed evaluations after 1 warm-up, "
                      f"{dt:.2f} s each; torch-autograd restatement of the reference path"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
