"""Second-moment decay that approaches 1 as training proceeds: ``beta2(t) = max(beta2_0, 1 - t^-c)``.

Counterpart of the reference's ``internlm/solver/schedulers/beta2_scheduler.py``; ``c <= 0`` keeps ``beta2_0`` for the whole run
(the shipped configs).  The value is a pure function of the iteration (``beta2_at``), so a resumed run only needs the step count.
"""
from typing import Optional


def beta2_at(iteration: int, init_beta2: float, c: float) -> float:
    if c <= 0 or iteration <= 0:
        return init_beta2
    return max(init_beta2, 1.0 - float(iteration) ** (-c))


class Beta2Scheduler:
    def __init__(self, optimizer, init_beta2, c=0.8, cur_iter=-1):
        self.optimizer, self.init_beta2, self.c = optimizer, init_beta2, c
        self.cur_iter = max(cur_iter, 0)      # -1 is the configs' "fresh run" mark

    def get_beta2(self) -> float:
        return beta2_at(self.cur_iter, self.init_beta2, self.c)

    def step(self, cur_iter: Optional[int] = None):
        """Advance (or jump to ``cur_iter``) and write the new beta2 into every parameter group."""
        self.cur_iter = cur_iter if cur_iter is not None else self.cur_iter + 1
        value = self.get_beta2()
        for group in self.optimizer.param_groups:
            group["betas"] = (group["betas"][0], value)
