"""The optimiser closure (tedeous/optimizers/closure.py:49-64): zero_grad, evaluate, backward."""


class Closure:
    def __init__(self, mixed_precision: bool, model):
        if mixed_precision:
            raise NotImplementedError('the fused path computes in fp32 / 3xTF32; AMP is not supported')
        self.model = model
        self.optimizer = model.optimizer
        self.normalized_loss_stop = model.normalized_loss_stop

    def _closure(self):
        self.optimizer.zero_grad()
        loss, loss_normalized = self.model.solution_cls.evaluate()
        loss.backward()
        self.model.cur_loss = loss_normalized if self.normalized_loss_stop else loss
        return loss

    def _closure_ngd(self):
        """tedeous/optimizers/closure.py:98-116: loss + gradient, then the residual fields and the loss function the
        natural-gradient step needs."""
        self.optimizer.zero_grad()
        sol = self.model.solution_cls
        loss, loss_normalized = sol.evaluate()
        loss.backward()
        self.model.cur_loss = loss_normalized if self.normalized_loss_stop else loss
        return sol.op, sol.bval, sol.true_bval, loss, sol.evaluate

    def get_closure(self, _type: str):
        if _type == 'NGD':
            return self._closure_ngd
        if _type in ('PSO', 'CSO', 'NNCG'):
            raise NotImplementedError(f'{_type} closure is not provided')
        return self._closure
