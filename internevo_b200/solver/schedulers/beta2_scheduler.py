"""beta2 schedule ``max(init_beta2, 1 - 1 / iter^c)`` (reference ``internlm/solver/schedulers/beta2_scheduler.py``)."""


class Beta2Scheduler:
    def __init__(self, optimizer, init_beta2, c=0.8, cur_iter=-1):
        self.cur_iter = 0 if cur_iter == -1 else cur_iter
        self.init_beta2 = init_beta2
        self.c = c
        self.optimizer = optimizer

    def step(self, cur_iter=None):
        self.cur_iter = self.cur_iter + 1 if cur_iter is None else cur_iter
        new_beta2 = self.get_beta2()
        for pg in self.optimizer.param_groups:
            beta1, _ = pg["betas"]
            pg["betas"] = (beta1, new_beta2)

    def get_beta2(self):
        if self.c <= 0 or self.cur_iter <= 0:
            return self.init_beta2
        return max(self.init_beta2, 1 - (1 / self.cur_iter**self.c))
