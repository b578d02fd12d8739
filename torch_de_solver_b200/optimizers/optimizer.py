"""Name -> optimiser map (tedeous/optimizers/optimizer.py:12-73).  Adam / AdamW / SGD / LBFGS / RMSprop are torch's (the
first three also have a fused CUDA-graph step, optimizers/fused.py); NGD runs on the per-residual Jacobian rows of the
fused path (optimizers/ngd.py).  PSO / CSO / NNCG are research optimisers outside the hot path (SURVEY 2
#13) and are not provided."""
from typing import Union

import torch
from torch.optim.lr_scheduler import CosineAnnealingWarmRestarts, ExponentialLR

from .ngd import NGD

_TORCH = {'Adam': torch.optim.Adam, 'AdamW': torch.optim.AdamW, 'SGD': torch.optim.SGD,
          'LBFGS': torch.optim.LBFGS, 'RMSprop': torch.optim.RMSprop, 'NGD': NGD}


class Optimizer:
    def __init__(self, optimizer: str, params: dict, gamma: Union[float, None] = None,
                 decay_every: Union[int, None] = None, cosine_scheduler_patience: Union[float, None] = None):
        self.optimizer = optimizer
        self.params = params
        self.gamma = gamma
        self.decay_every = decay_every
        self.cosine_scheduler_patience = cosine_scheduler_patience

    def optimizer_choice(self, mode, model):
        if self.optimizer not in _TORCH:
            raise NotImplementedError(f'optimizer {self.optimizer!r} is not provided (available: {sorted(_TORCH)})')
        cls = _TORCH[self.optimizer]
        if mode in ('NN', 'autograd'):
            optimizer = cls(model.parameters(), **self.params)
        else:
            optimizer = cls([model.requires_grad_()], **self.params)
        if self.gamma is not None:
            self.scheduler = ExponentialLR(optimizer, gamma=self.gamma)
        if self.cosine_scheduler_patience is not None:
            self.scheduler = CosineAnnealingWarmRestarts(optimizer, self.cosine_scheduler_patience)
        return optimizer
