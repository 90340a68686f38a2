"""Fused trainer: the reference's `class pinn` (train_newmethod.py:46-209; the train.py:46-200
two-point-set form through `fid_*`) on top of one fused evaluation per loss call.

Same config keys, same schedule (Adam + StepLR for adam_optimizer.max_it
iterations, then ONE `LBFGS.step(closure)`), same `log.txt` CSV (`Epoch, Fidelity Loss, Residual
Loss, Total Loss`, one line per loss evaluation, `%.5e`) and the same checkpoint cadence -- but
the three logged numbers stay on the device in a ring and are flushed to the host every
`log_every` evaluations instead of forcing three syncs per call.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .dnn import DNN
from .fused import JetLoss
from .lbfgs import LBFGS, FusedAdam
from .spec import DIR_ORDER, FIELD_ORDER, PassSpec


class FusedClosure:
    """Callable with the torch closure contract *and* the flat fast path LBFGS / LBFGSBOptimizer
    look for (`flat_loss_and_grad`)."""

    def __init__(self, owner):
        self.owner = owner

    def flat_loss_and_grad(self, flat_params, flat_grad):
        return self.owner._evaluate(flat_params, flat_grad)

    def __call__(self):
        o = self.owner
        parts = o._evaluate(o.flat, o.flat_grad)
        for p, g in zip(o.dnn.parameters(), o.dnn.split_flat(o.flat_grad)):
            p.grad = g
        return parts[2]


class pinn:
    def __init__(self, config, data_input, data_true, residual="continuity_only",
                 fid_input=None, fid_true=None, device=None, log_dir=None, group=None,
                 log_every=1000, precision="fp32", dump_at=50000, dump_path='data_at50k.mat'):
        """data_input [N,d] / data_true [N,n_true]: numpy arrays exactly as the reference's
        __main__ builds them (train_newmethod.py:226-255), or device tensors from data.assemble_points.
        When fid_input/fid_true are given the train.py form is used: data misfit on (fid_input, fid_true),
        residual on data_input.  dump_at / dump_path: the prediction dump of train_newmethod.py:141-153
        (`data_at50k.mat`, written by the evaluation that starts with iter == 50000)."""
        self.config = config
        lay = config['layers']
        self.layers = [lay['input_features']] + [lay['hidden_width']] * lay['hidden_layers'] \
            + [lay['output_features']]
        self.dropout_rate = lay.get('dropout_rate', 0.0)
        self.init_type = lay.get('init_type', 'xavier')
        self.device = torch.device(device if device is not None else "cuda")
        self.dnn = DNN(self.layers, self.dropout_rate, self.init_type).to(self.device)
        self.flat = self.dnn.flat_params()
        self.flat_grad = torch.zeros_like(self.flat)
        self.adam_maxit = int(config['adam_optimizer']['max_it'])
        self.iter = 0
        self.weight_fidelity = config['loss']['weight_fid_loss']
        self.weight_residual = config['loss']['weight_res_loss']
        self.log_dir = log_dir
        self.log_every = int(log_every)
        self._ring = torch.zeros(self.log_every, 4, dtype=torch.float32, device=self.device)
        self._ring_n = 0
        self.history = []                 # flushed (iter, fid, res, total)

        self.dump_at, self.dump_path, self.group = dump_at, dump_path, group
        names_d, names_f = DIR_ORDER[residual], FIELD_ORDER[residual]

        def t(a):
            if isinstance(a, torch.Tensor):
                return a.to(device=self.device, dtype=torch.float32).contiguous()
            return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).to(self.device)
        if fid_input is None:
            # train_newmethod.py form: inputs in config['data']['inputs'] order, outputs =
            # trues followed by unknowns (train_newmethod.py:136-139)
            in_names = list(config['data']['inputs'].keys())
            out_names = list(config['data']['trues']) + list(config['data']['unknowns'])
            n_true = len(config['data']['trues'])
            spec = PassSpec(layers=self.layers, activation=self.dnn.activation_name, kind=residual,
                            dirs={n: in_names.index(n) for n in names_d},
                            fields={n: out_names.index(n) for n in names_f},
                            target_cols=list(range(n_true)), w_fid=self.weight_fidelity,
                            w_res=self.weight_residual, precision=precision)
            self.jl = JetLoss(spec, t(data_input), t(data_true), group=group)
            self.out_names = out_names
        else:
            in_names = list(config['data_residual']['inputs'].keys())
            out_names = list(config['data_residual']['outputs'])
            fid_out = list(config['data_fidelity']['outputs'])
            # per-output weights (train.py:94-95); config.json / config_txyz.json do not carry them: 1
            tw = [config['loss'].get(f'weight_{k}_loss', 1) for k in fid_out]
            self.out_names = out_names
            common = dict(layers=self.layers, activation=self.dnn.activation_name,
                          w_fid=self.weight_fidelity, w_res=self.weight_residual, precision=precision)
            sres = PassSpec(kind=residual, dirs={n: in_names.index(n) for n in names_d},
                            fields={n: out_names.index(n) for n in names_f}, **common)
            sfid = PassSpec(kind="none", target_cols=[out_names.index(k) for k in fid_out],
                            target_w=tw, **common)
            self.jl = JetLoss(sres, t(data_input), None, fid=(sfid, t(fid_input), t(fid_true)),
                              group=group)
        self.init_optimizers()

    # train_newmethod.py:92-117
    def init_optimizers(self):
        a, l = self.config['adam_optimizer'], self.config['lbfgs_optimizer']
        self.optimizer_Adam = FusedAdam(self.dnn.parameters(), lr=a['learning_rate'])
        self.scheduler_Adam = torch.optim.lr_scheduler.StepLR(
            self.optimizer_Adam, step_size=a['scheduler_step_size'], gamma=a['scheduler_gamma'])
        self.optimizer_LBFGS = LBFGS(
            self.dnn.parameters(), lr=l['learning_rate'], max_iter=l['max_it'],
            max_eval=l['max_evaluation'], history_size=l['history_size'],
            tolerance_grad=l['tolerance_grad'], tolerance_change=l['tolerance_change'],
            line_search_fn=l['line_search_fn'])

    # ---- one evaluation: loss parts on device + flat gradient; bookkeeping of loss_func ----------
    def predictions(self, flat=None):
        """Network outputs [N, o] at the residual points (this rank's shard), value-only fused forward."""
        flat = self.flat if flat is None else flat
        out = torch.empty(self.jl.res.n, self.layers[-1], dtype=torch.float32, device=self.device)
        self.jl.loss(flat, out=out)
        return out

    def dump_predictions(self, path, flat=None):
        """train_newmethod.py:141-153: {'pred_<name>': [N,1] float32} for the true and unknown variables -> .mat"""
        out = self.predictions(flat)
        if self.group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(self.group)
            sizes = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([out.shape[0]], dtype=torch.int64, device=self.device), group=self.group)
            nmax = int(max(s.item() for s in sizes))
            pad = torch.zeros(nmax, out.shape[1], device=self.device)
            pad[:out.shape[0]] = out
            allp = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(allp, pad, group=self.group)
            out = torch.cat([a[:int(s.item())] for a, s in zip(allp, sizes)])
            if dist.get_rank(self.group) != 0:
                return None
        import scipy.io as sio
        host = out.cpu().numpy()
        data = {f'pred_{k}': host[:, i:i + 1].copy() for i, k in enumerate(self.out_names)}
        sio.savemat(path, data)
        print(f'Data saved to {path} after {self.iter:,} iterations.')
        return data

    def _evaluate(self, flat, grad):
        if self.dump_at is not None and self.iter == self.dump_at:   # before the counter moves, like the reference
            self.dump_predictions(self.dump_path, flat)
        parts = self.jl.loss_and_grad(flat, grad)
        self.iter += 1
        self._ring[self._ring_n].copy_(parts)
        self._ring[self._ring_n, 3] = float(self.iter)
        self._ring_n += 1
        if self._ring_n == self.log_every:
            self.flush_log()
        if self.log_dir is not None:
            it = self.iter
            if (it <= 45000 and it % 10000 == 0) or (it > 45000 and it % 1000 == 0):
                torch.save(self.dnn, os.path.join(self.log_dir, f'model_{it}.pth'))
        return parts

    def loss_func(self):
        """Reference-shaped entry (train_newmethod.py:120): returns the total loss tensor and leaves
        the flat gradient in self.flat_grad (the caller does not need loss.backward())."""
        return self._evaluate(self.flat, self.flat_grad)[2]

    def flush_log(self):
        if self._ring_n == 0:
            return
        rows = self._ring[:self._ring_n].cpu().numpy()
        self._ring_n = 0
        for fid, res, tot, it in rows:
            self.history.append((int(it), float(fid), float(res), float(tot)))
        if self.log_dir is not None:
            path = os.path.join(self.log_dir, 'log.txt')
            new = not os.path.exists(path) or os.stat(path).st_size == 0
            with open(path, 'a') as f:
                if new:
                    f.write('Epoch, Fidelity Loss, Residual Loss, Total Loss\n')
                for fid, res, tot, it in rows:
                    f.write(f'{int(it)}, {fid:.5e}, {res:.5e}, {tot:.5e}\n')
        fid, res, tot, it = rows[-1]
        if int(it) % 1000 == 0:
            print(f'Epoch {int(it)}, Fidelity Loss: {fid:.5e}, Residual Loss: {res:.5e}, '
                  f'Total Loss: {tot:.5e}')

    # train_newmethod.py:194-209
    def train(self):
        self.dnn.train()
        for _ in range(self.adam_maxit):
            self._evaluate(self.flat, self.flat_grad)
            self.optimizer_Adam.step(flat_grad=self.flat_grad)
            self.scheduler_Adam.step()
        self.optimizer_LBFGS.step(FusedClosure(self))
        self.flush_log()
