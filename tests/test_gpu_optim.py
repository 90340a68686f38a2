"""GPU tests of the optimiser kernels and the device-backed L-BFGS / Adam / trainer against torch's own
implementations (torch.optim.LBFGS / Adam are the third-party code the reference calls,
train_newmethod.py:95-117)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import autograd_port as ap
from oracle import jet_oracle as jo
from tests import cases

pytestmark = pytest.mark.gpu


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("P,hist,m,head", [(1000, 5, 0, 0), (4097, 7, 3, 0), (41703, 11, 11, 4),
                                           (462852, 9, 6, 7), (33, 4, 4, 3)])
def test_two_loop_direction_matches_numpy(P, hist, m, head):
    from pinn_depthestimation_b200 import _cabi
    rs = np.random.RandomState(P + m)
    S = rs.standard_normal((hist, P)).astype(np.float32) * 0.1
    Y = (S + 0.05 * rs.standard_normal((hist, P))).astype(np.float32)
    g = rs.standard_normal(P).astype(np.float32)
    rho = np.zeros(hist, np.float32)
    for i in range(hist):
        rho[i] = 1.0 / float(Y[i].astype(np.float64) @ S[i].astype(np.float64))
    hd = np.float32(0.7)
    # numpy two-loop in float64 (torch/optim/lbfgs.py:432-447)
    q = -g.astype(np.float64)
    al = np.zeros(m)
    slots = [(head + i) % hist for i in range(m)]
    for i in range(m - 1, -1, -1):
        al[i] = S[slots[i]].astype(np.float64) @ q * rho[slots[i]]
        q -= al[i] * Y[slots[i]]
    r = q * float(hd)
    for i in range(m):
        be = Y[slots[i]].astype(np.float64) @ r * rho[slots[i]]
        r += (al[i] - be) * S[slots[i]]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    d = torch.empty(P, device=dev)
    scratch = torch.zeros(2 * hist + 64, device=dev)
    tS, tY, trho, tg, thd = t(S), t(Y), t(rho), t(g), t(np.array([hd]))
    _cabi.check(_cabi.lib().pinn_lbfgs_direction(
        _cabi.ptr(tS), _cabi.ptr(tY), _cabi.ptr(trho), _cabi.ptr(thd), _cabi.ptr(tg), _cabi.ptr(d),
        hist, m, head, P, _cabi.ptr(scratch), _st()))
    got = d.cpu().numpy().astype(np.float64)
    assert np.linalg.norm(got - r) <= 2e-5 * np.linalg.norm(r)
    # deterministic: a second launch gives the same bits
    d2 = torch.empty(P, device=dev)
    _cabi.check(_cabi.lib().pinn_lbfgs_direction(
        _cabi.ptr(tS), _cabi.ptr(tY), _cabi.ptr(trho), _cabi.ptr(thd), _cabi.ptr(tg), _cabi.ptr(d2),
        hist, m, head, P, _cabi.ptr(scratch), _st()))
    assert torch.equal(d, d2)


def test_vec_stats_and_axpy():
    from pinn_depthestimation_b200 import _cabi
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    for n in (1, 31, 4096, 462852):
        a = rs.standard_normal(n).astype(np.float32)
        b = rs.standard_normal(n).astype(np.float32)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        out = torch.zeros(8, device=dev)
        _cabi.check(_cabi.lib().pinn_vec_stats(_cabi.ptr(ta), _cabi.ptr(tb), n, _cabi.ptr(out), _st()))
        o = out.cpu().numpy()
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        ref = [a64 @ b64, np.abs(a64).sum(), np.abs(a64).max(), np.abs(b64).max(), a64 @ a64, b64 @ b64]
        scale = [np.sqrt((a64 @ a64) * (b64 @ b64)), ref[1], ref[2], ref[3], ref[4], ref[5]]
        for k in range(6):
            assert abs(o[k] - ref[k]) <= 2e-6 * max(scale[k], 1e-30), (n, k)
        y = tb.clone()
        _cabi.check(_cabi.lib().pinn_axpy(0.37, _cabi.ptr(ta), _cabi.ptr(y), n, _st()))
        np.testing.assert_allclose(y.cpu().numpy(), b + np.float32(0.37) * a, rtol=1e-6, atol=1e-6)


def test_fused_adam_matches_torch_adam_with_steplr():
    from pinn_depthestimation_b200.lbfgs import FusedAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    shapes = [(20, 2), (20,), (20, 20), (20,), (3, 20), (3,)]
    ref = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-2)
    o_mine = FusedAdam(mine, lr=1e-2)
    s_ref = torch.optim.lr_scheduler.StepLR(o_ref, step_size=5, gamma=0.8)
    s_mine = torch.optim.lr_scheduler.StepLR(o_mine, step_size=5, gamma=0.8)
    for it in range(23):
        gs = [torch.randn(s, device=dev) * (1 + it) for s in shapes]
        for p, q, g in zip(ref, mine, gs):
            p.grad, q.grad = g.clone(), g.clone()
        o_ref.step(); s_ref.step()
        o_mine.step(); s_mine.step()
    for p, q in zip(ref, mine):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6)


def _setup_problem(name, dev):
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    case, _ = cases.load(name)
    sres, _ = pass_specs(case)
    flat, X, T, _, _ = cases.data(case, np.float32)
    jl = JetLoss(sres, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
    return case, flat, X, T, jl


def _torch_reference_lbfgs(case, flat, X, T, **kw):
    """torch.optim.LBFGS driving the reference algorithm (autograd port) on CPU."""
    ospec, _ = cases.specs(case)
    p = torch.nn.Parameter(torch.from_numpy(flat.copy()))
    opt = torch.optim.LBFGS([p], **kw)
    Xt, Tt = torch.from_numpy(X), torch.from_numpy(T)
    losses = []

    def closure():
        opt.zero_grad()
        r = ap.loss_and_grad(ospec, p.detach(), Xt, Tt)
        p.grad = r["grad"].clone()
        losses.append(float(r["loss"]))
        return r["loss"]
    opt.step(closure)
    return opt.state[p]["n_iter"], opt.state[p]["func_evals"], losses, p.detach().numpy()


@pytest.mark.parametrize("ls", ["strong_wolfe", None])
def test_lbfgs_training_curve_matches_torch_lbfgs(ls):
    """Same closure contract, same hyper-parameters as train_newmethod.py:108-117; the loss-vs-evaluation
    curve must stay within a narrow envelope of torch.optim.LBFGS on the reference algorithm."""
    from pinn_depthestimation_b200.lbfgs import LBFGS
    dev = torch.device("cuda:0")
    case, flat, X, T, jl = _setup_problem("cmb_h_small", dev)
    kw = dict(lr=1 if ls else 0.05, max_iter=25, max_eval=40, history_size=100, tolerance_grad=1e-5,
              tolerance_change=1e-7, line_search_fn=ls)
    n_ref, ev_ref, l_ref, _ = _torch_reference_lbfgs(case, flat, X, T, **kw)

    p = torch.nn.Parameter(torch.from_numpy(flat.copy()).to(dev))
    opt = LBFGS([p], **kw)
    mine = []

    class Closure:
        def flat_loss_and_grad(self, fp, fg):
            parts = jl.loss_and_grad(fp, fg)
            mine.append(parts[2].item())
            return parts
    opt.step(Closure())
    st = opt.state[p]
    assert st["n_iter"] == n_ref and st["func_evals"] == ev_ref, (st["n_iter"], n_ref, st["func_evals"], ev_ref)
    assert len(mine) == len(l_ref)
    l_ref, mine = np.array(l_ref), np.array(mine)
    assert abs(mine[0] - l_ref[0]) <= 1e-5 * l_ref[0]
    assert np.max(np.abs(mine - l_ref) / l_ref) <= 2e-3        # envelope over the whole curve
    assert mine[-1] < 0.8 * mine[0]


def test_lbfgs_accepts_the_reference_closure_contract():
    """closure = zero_grad -> loss_func -> backward -> return loss (train_newmethod.py:204-208) on the
    drop-in modules; history carries over between step() calls like torch's."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
    import dnn, physics
    from pinn_depthestimation_b200.lbfgs import LBFGS
    dev = torch.device("cuda:0")
    case, _ = cases.load("cmb_h_small")
    flat, X, T, _, _ = cases.data(case, np.float32)
    m = dnn.DNN(case["layers"], 0.0, "xavier").to(dev)
    with torch.no_grad():
        m.flat_params().copy_(torch.from_numpy(flat).to(dev))
    x = torch.tensor(X[:, 0:1].astype(np.float64), requires_grad=True).float().to(dev)
    y = torch.tensor(X[:, 1:2].astype(np.float64), requires_grad=True).float().to(dev)
    Tt = torch.from_numpy(T).to(dev)
    opt = LBFGS(m.parameters(), lr=1, max_iter=6, max_eval=10, history_size=100, tolerance_grad=1e-5,
                tolerance_change=1e-7, line_search_fn="strong_wolfe")
    seen = []

    def closure():
        opt.zero_grad()
        pred = m(torch.cat([x, y], dim=-1))
        fid = sum(torch.nn.functional.mse_loss(pred[:, i:i + 1], Tt[:, i:i + 1]) for i in range(2))
        res = physics.continuity_only(x, y, pred[:, 2:3], pred[:, 0:1], pred[:, 1:2])
        loss = fid + res
        loss.backward()
        seen.append(loss.item())
        return loss
    opt.step(closure)
    first = list(seen)
    opt.step(closure)
    assert min(seen) < 0.7 * first[0]
    assert opt.state[list(m.parameters())[0]]["n_iter"] == 12


def test_adam_phase_curve_matches_reference_algorithm():
    """Training-curve equivalence for the Adam phase (train_newmethod.py:197-202): 150 iterations of
    torch.optim.Adam on the reference algorithm (CPU) vs FusedAdam on the fused kernel."""
    from pinn_depthestimation_b200.lbfgs import FusedAdam
    dev = torch.device("cuda:0")
    case, flat, X, T, jl = _setup_problem("cmb_h_small", dev)
    ospec, _ = cases.specs(case)
    p = torch.nn.Parameter(torch.from_numpy(flat.copy()))
    o = torch.optim.Adam([p], lr=1e-3)
    Xt, Tt = torch.from_numpy(X), torch.from_numpy(T)
    ref = []
    for _ in range(150):
        r = ap.loss_and_grad(ospec, p.detach(), Xt, Tt)
        p.grad = r["grad"]
        ref.append(float(r["loss"]))
        o.step()
    q = torch.nn.Parameter(torch.from_numpy(flat.copy()).to(dev))
    om = FusedAdam([q], lr=1e-3)
    g = torch.empty_like(q)
    mine = torch.zeros(150, device=dev)
    for i in range(150):
        parts = jl.loss_and_grad(q.detach(), g)
        mine[i] = parts[2]
        om.step(flat_grad=g)
    mine = mine.cpu().numpy()
    ref = np.array(ref)
    assert np.max(np.abs(mine - ref) / ref) <= 1e-3
    assert mine[-1] < mine[0]


def test_trainer_runs_the_reference_schedule_and_log_format(tmp_path):
    from pinn_depthestimation_b200.trainer import pinn
    config = {
        "layers": {"input_features": 2, "hidden_layers": 6, "hidden_width": 20, "output_features": 3,
                   "dropout_rate": 0.0, "init_type": "xavier"},
        "adam_optimizer": {"max_it": 30, "learning_rate": 1e-3, "scheduler_step_size": 10,
                           "scheduler_gamma": 0.8},
        "lbfgs_optimizer": {"max_it": 8, "learning_rate": 1, "max_evaluation": 12.0, "history_size": 100,
                            "tolerance_grad": 1e-5, "tolerance_change": 1e-7, "line_search_fn": "strong_wolfe"},
        "loss": {"weight_fid_loss": 1, "weight_res_loss": 1},
        "data": {"inputs": {"x": {"requires_grad": ["true"]}, "y": {"requires_grad": ["true"]}},
                 "trues": ["U", "V"], "unknowns": ["h"]},
    }
    X, T = jo.make_points(400, 2, 2, seed=3)
    torch.manual_seed(1234)
    model = pinn(config, X, T, log_dir=str(tmp_path), log_every=16)
    model.train()
    lines = open(tmp_path / "log.txt").read().strip().splitlines()
    assert lines[0] == "Epoch, Fidelity Loss, Residual Loss, Total Loss"
    n_evals = model.iter
    assert len(lines) == 1 + n_evals and n_evals >= 31
    first = [float(v) for v in lines[1].split(",")]
    last = [float(v) for v in lines[-1].split(",")]
    assert int(first[0]) == 1 and int(last[0]) == n_evals
    assert abs(first[1] + first[2] - first[3]) <= 1e-4 * first[3]
    assert min(float(l.split(",")[3]) for l in lines[1:]) < first[3]
    torch.save(model.dnn, tmp_path / "model.pth")
    again = torch.load(tmp_path / "model.pth", weights_only=False)
    assert torch.equal(again.flat_params(), model.dnn.flat_params())


def test_train_main_script_runs_a_reference_style_config(tmp_path):
    """python -m pinn_depthestimation_b200.train_main on a config with the reference's key layout
    (config_CMB_h.json), synthetic points standing in for the unshipped .mat file."""
    import json
    from pinn_depthestimation_b200 import train_main
    cfg = {
        "layers": {"input_features": 2, "hidden_layers": 5, "hidden_width": 20, "output_features": 3,
                   "dropout_rate": 0.0, "init_type": "xavier"},
        "adam_optimizer": {"max_it": 20, "learning_rate": 1e-3, "scheduler_step_size": 10000,
                           "scheduler_gamma": 0.8},
        "lbfgs_optimizer": {"max_it": 5, "learning_rate": 1, "max_evaluation": 8, "history_size": 100,
                            "tolerance_grad": 1e-5, "tolerance_change": 1e-7, "line_search_fn": "strong_wolfe"},
        "loss": {"weight_fid_loss": 1, "weight_res_loss": 1},
        "data": {"file": "unused.mat", "inputs": {"x": {"requires_grad": ["true"]}, "y": {"requires_grad": ["true"]}},
                 "trues": ["U", "V"], "unknowns": ["h"]},
        "data_test": {"x_min": 25.0, "x_max": 33.0, "y_min": -13.0, "y_max": 13.0},
    }
    path = tmp_path / "config.json"
    path.write_text(json.dumps(cfg))
    model = train_main.main(["--config", str(path), "--synthetic", "300", "--log-dir", str(tmp_path / "log")])
    assert (tmp_path / "log" / "model.pth").exists()
    assert (tmp_path / "log" / "log.txt").exists()
    assert model.history[-1][3] < model.history[0][3]


def _run_both_line_searches(name, max_iter, max_eval, steps=2):
    from pinn_depthestimation_b200.lbfgs import LBFGS
    dev = torch.device("cuda:0")
    case, flat, X, T, jl = _setup_problem(name, dev)
    kw = dict(lr=1, max_iter=max_iter, max_eval=max_eval, history_size=7, tolerance_grad=0.0, tolerance_change=0.0,
              line_search_fn="strong_wolfe")
    runs = {}
    for device_ls in (False, True):
        p = torch.nn.Parameter(torch.from_numpy(flat.copy()).to(dev))
        opt = LBFGS([p], **kw)
        opt.device_line_search = device_ls
        curve = []

        class Closure:
            def flat_loss_and_grad(self, fp, fg):
                parts = jl.loss_and_grad(fp, fg)
                curve.append(parts[2].item())
                return parts
        first = opt.step(Closure())
        n1 = (opt.state[p]["n_iter"], opt.state[p]["func_evals"])
        for _ in range(steps - 1):
            opt.step(Closure())      # history and the step length carry over (torch keeps them in self.state)
        st = opt.state[p]
        runs[device_ls] = (float(first), n1, st["n_iter"], st["func_evals"], np.array(curve))
    return runs[False], runs[True]


@pytest.mark.parametrize("name", ["cmb_h_small", "txyz", "ftemp_small", "ragged"])
def test_device_line_search_takes_the_same_branches_as_the_host_restatement(name):
    """The device state machine (csrc/lbfgs_dev.cu) against the host restatement of torch's step / _strong_wolfe
    (LBFGS._step_host, itself held to torch's private functions on CPU) over the first iterations, where both are still
    on the same path: identical iteration and evaluation counts, the same loss at every evaluation."""
    h, d = _run_both_line_searches(name, 8, 12, steps=1)
    print(f"{name}: host {h[1]}, device {d[1]}; loss {h[4][0]:.4e} -> {h[4][-1]:.4e}")
    assert d[0] == h[0] == h[4][0]           # step() returns the FIRST evaluation's loss (torch's orig_loss)
    assert d[1] == h[1]
    assert len(d[4]) == len(h[4])
    assert np.max(np.abs(d[4] - h[4]) / np.abs(h[4])) <= 1e-3
    assert h[4][-1] < h[4][0]


@pytest.mark.parametrize("name,max_iter,max_eval", [("cmb_h_small", 60, 75), ("txyz", 40, 50), ("ftemp_small", 80, 100),
                                                    ("cmb_h_small", 30, 17)])
def test_device_line_search_long_runs_stay_in_the_host_envelope(name, max_iter, max_eval):
    """Long runs over bracket, zoom and max_eval exits and two step() calls.  The device path forms the direction in
    coefficient space (double-precision Gram algebra), the host path with the FP32 vector two-loop: mathematically the
    same d, rounded differently, and the gradient kernel's float atomics add 1e-7 noise of their own -- strong Wolfe is
    discontinuous in its inputs, so late bracket decisions flip and the two runs are compared as an envelope."""
    h, d = _run_both_line_searches(name, max_iter, max_eval, steps=2)
    print(f"{name}: host {h[1]} -> ({h[2]}, {h[3]}), device {d[1]} -> ({d[2]}, {d[3]}); loss {h[4][0]:.4e} -> {h[4][-1]:.4e} / {d[4][-1]:.4e}")
    assert d[0] == h[0]
    assert abs(d[2] - h[2]) <= max(1, h[2] // 10)
    assert abs(d[3] - h[3]) <= max(3, h[3] // 5)
    k = min(8, len(h[4]), len(d[4]))
    assert np.max(np.abs(d[4][:k] - h[4][:k]) / np.abs(h[4][:k])) <= 1e-3
    assert d[4][-1] < 0.5 * d[4][0] and h[4][-1] < 0.5 * h[4][0]
    assert 0.5 <= min(d[4]) / min(h[4]) <= 2.0
