"""Learning-rate schedule of the reference (learning_rates.py:22-88, built by pretrain_gpt2.py:161-178): linear
warm-up to `start_lr`, then linear / cosine decay or a constant.  Same constructor arguments and the same
`state_dict` keys (warmup_iter, num_iters, decay_style, end_iter, decay_ratio), so the 'lr_scheduler' entry of a
reference checkpoint restores the schedule of a resumed run (utils.py:353-356)."""
import math


class AnnealingLR:
    DECAY_STYLES = ('linear', 'cosine', 'exponential', 'constant', 'None')

    def __init__(self, optimizer, start_lr, warmup_iter, num_iters, decay_style=None, last_iter=-1, decay_ratio=0.5):
        if warmup_iter > num_iters:
            raise AssertionError('warmup_iter must not exceed num_iters')
        self.optimizer = optimizer
        self.start_lr = start_lr
        self.warmup_iter = warmup_iter
        self.end_iter = num_iters
        self.decay_style = decay_style.lower() if isinstance(decay_style, str) else None
        self.decay_ratio = 1 / decay_ratio          # the reference stores the inverse (learning_rates.py:37)
        self.num_iters = 0
        self.step(last_iter + 1)

    def get_lr(self):
        it, warm = self.num_iters, self.warmup_iter
        if warm > 0 and it <= warm:                 # linear warm-up (learning_rates.py:44-45)
            return float(self.start_lr) * it / warm
        if self.decay_style == 'linear':
            return self.start_lr * ((self.end_iter - (it - warm)) / self.end_iter)
        if self.decay_style == 'cosine':
            frac = min(1.0, (it - warm) / self.end_iter)
            r = self.decay_ratio
            return self.start_lr / r * ((math.cos(math.pi * frac) + 1) * (r - 1) / 2 + 1)
        return self.start_lr                        # 'exponential' is unimplemented in the reference too (:53-55)

    def step(self, step_num=None):
        self.num_iters = self.num_iters + 1 if step_num is None else step_num
        lr = self.get_lr()
        for group in self.optimizer.param_groups:
            group['lr'] = lr

    def state_dict(self):
        return {'warmup_iter': self.warmup_iter, 'num_iters': self.num_iters, 'decay_style': self.decay_style,
                'end_iter': self.end_iter, 'decay_ratio': self.decay_ratio}

    def load_state_dict(self, sd):
        # like the reference (:79-88): start_lr and end_iter come from the command line of the resumed run
        self.warmup_iter = sd['warmup_iter']
        self.decay_style = sd['decay_style']
        if 'decay_ratio' in sd:
            self.decay_ratio = sd['decay_ratio']
        self.step(sd['num_iters'])
