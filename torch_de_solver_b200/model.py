"""`Model.compile(mode, ...)` / `Model.train(optimizer, epochs, ...)` (tedeous/model.py:25-195): the same driver
loop around `Solution.evaluate`, with this repo's fused Solution underneath."""
import datetime
from typing import List, Union

import torch

from .data import Conditions, Domain, Equation
from .input_preprocessing import Operator_bcond_preproc
from .optimizers.closure import Closure
from .optimizers.optimizer import Optimizer
from .solution import Solution


class _Callbacks:
    """Duck-typed Keras-style hook list (tedeous/callbacks/callback_list.py)."""

    def __init__(self, callbacks, model):
        self.callbacks = list(callbacks or [])
        for cb in self.callbacks:
            if hasattr(cb, 'set_model'):
                cb.set_model(model)
            else:
                cb.model = model

    def _call(self, name, *a):
        for cb in self.callbacks:
            fn = getattr(cb, name, None)
            if fn is not None:
                fn(*a)

    def on_train_begin(self): self._call('on_train_begin')
    def on_train_end(self): self._call('on_train_end')
    def on_epoch_begin(self): self._call('on_epoch_begin')
    def on_epoch_end(self): self._call('on_epoch_end')


class Model:
    def __init__(self, net: Union[torch.nn.Module, torch.Tensor], domain: Domain, equation: Equation,
                 conditions: Conditions, batch_size: int = None):
        self.net = net
        self.domain = domain
        self.equation = equation
        self.conditions = conditions
        self.batch_size = batch_size
        self._check = None

    def compile(self, mode: str, lambda_operator, lambda_bound, normalized_loss_stop: bool = False,
                h: float = 0.001, inner_order: str = '1', boundary_order: str = '2',
                derivative_points: int = 2, weak_form: List[callable] = None, tol: float = 0,
                removed_domains: list = None, **fused_options):
        """`fused_options` (extensions): shard=(rank, world), process_group, nn_interior='jet'|'literal', impl."""
        self.mode = mode
        self.lambda_bound = lambda_bound
        self.lambda_operator = lambda_operator
        self.normalized_loss_stop = normalized_loss_stop
        self.weak_form = weak_form
        self.removed_domains = removed_domains
        grid = self.domain.build(mode=mode, removed_domains=removed_domains)
        if isinstance(self.net, torch.nn.Module):
            self.net.to(grid.dtype)
        bconds = self.conditions.build(self.domain.variable_dict)
        self.equation_cls = Operator_bcond_preproc(grid, self.equation.equation_lst, bconds, h=h,
                                                   inner_order=inner_order,
                                                   boundary_order=boundary_order).set_strategy(mode)
        if self.batch_size is not None and len(grid) < self.batch_size:
            self.batch_size = None
        self.solution_cls = Solution(grid, self.equation_cls, self.net, mode, weak_form, lambda_operator,
                                     lambda_bound, tol, derivative_points, batch_size=self.batch_size,
                                     **fused_options)

    def _fused_train_step(self, optimizer: Optimizer, mixed_precision: bool):
        """The graph-captured training step (optimizers/fused.py) when the optimiser and the loss allow it: Adam / AdamW /
        SGD without cosine restarts, default loss (no causal weights, no weak form).  TDB200_EAGER_TRAIN=1 keeps the eager
        loop (torch optimiser + closure), which is also what every other configuration uses."""
        import os
        from .optimizers.fused import FusedOptimizer, TrainStep, _KIND
        sol = self.solution_cls
        if (optimizer.optimizer not in _KIND or mixed_precision or optimizer.cosine_scheduler_patience is not None
                or sol.tol != 0 or sol.weak_form not in (None, []) or os.environ.get('TDB200_EAGER_TRAIN')
                or getattr(sol, '_callable_coeffs', 'once') != 'once' or getattr(sol, '_batching', False)
                or getattr(sol, '_hybrid', False)):
            return None
        params = [sol.model] if sol.mode == 'mat' else sol._ir.net.param_tensors()
        try:
            opt = FusedOptimizer(optimizer.optimizer, params, **optimizer.params)
            return TrainStep(sol, opt)
        except (NotImplementedError, ValueError):
            return None

    def train(self, optimizer: Optimizer, epochs: int, info_string_every: Union[int, None] = None,
              mixed_precision: bool = False, save_model: bool = False, model_name: Union[str, None] = None,
              callbacks: Union[List, None] = None):
        self.t = 1
        self.stop_training = False
        callbacks = _Callbacks(callbacks, self)
        callbacks.on_train_begin()
        self.net = self.solution_cls.model
        self.optimizer = optimizer.optimizer_choice(self.mode, self.net)
        closure = Closure(mixed_precision, self).get_closure(optimizer.optimizer)
        self.min_loss, _ = self.solution_cls.evaluate()
        self.cur_loss = self.min_loss
        print('[{}] initial (min) loss is {}'.format(datetime.datetime.now(), self.min_loss.item()))
        fused = self._fused_train_step(optimizer, mixed_precision)
        lr0, decays = (fused.opt.lr, 0) if fused is not None else (None, 0)
        while fused is not None and self.t < epochs and self.stop_training is False:
            # one CUDA-graph replay per epoch: pack -> kernels -> reduce -> [all-reduce] -> parameter update; the loss
            # stays on the device (callbacks that read model.cur_loss synchronise, as they do in the reference)
            callbacks.on_epoch_begin()
            out = fused.step()
            self.cur_loss = out[1:2] if self.normalized_loss_stop else out[0:1]
            self.solution_cls.loss, self.solution_cls.loss_normalized = out[0:1], out[1:2]
            self.solution_cls._last_out = out
            if optimizer.gamma is not None and self.t % optimizer.decay_every == 0:
                decays += 1
                fused.opt.set_lr(lr0 * optimizer.gamma ** decays)
            callbacks.on_epoch_end()
            self.t += 1
        while fused is None and self.t < epochs and self.stop_training is False:
            callbacks.on_epoch_begin()
            self.optimizer.zero_grad()
            for _ in range(self.solution_cls.operator.n_batches):       # one optimiser step per mini-batch (model.py:178-184)
                self.optimizer.step(closure)
                if optimizer.gamma is not None and self.t % optimizer.decay_every == 0:
                    optimizer.scheduler.step()
            callbacks.on_epoch_end()
            self.t += 1
        callbacks.on_train_end()
        if save_model:
            torch.save({'model': self.net}, (model_name or 'tedeous_b200_model') + '.tar')
