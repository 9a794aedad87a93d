"""Exponential moving average of parameters -- drop-in for the reference's `models.ema`
(/root/reference/Reflected-Diffusion/models/ema.py:11-101): same constructor, `update`, `copy_to`,
`store`, `restore`, `state_dict`, `load_state_dict` and the same checkpoint keys
(`decay`, `num_updates`, `shadow_params`), so a reference checkpoint's `ema` entry loads unchanged.
The sampler caller swaps the averaged weights in around every sampling call
(Benchmark/gto_halo_benchmarking.py:230-239); the B200 NCSNpp notices the swap and re-packs.
Batched `torch._foreach_*` updates instead of a Python loop per tensor."""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if not 0.0 <= decay <= 1.0:
            raise ValueError('Decay must be between 0 and 1')
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.shadow_params = [p.detach().clone() for p in parameters if p.requires_grad]
        self.collected_params = []

    @staticmethod
    def _trainable(parameters):
        return [p for p in parameters if p.requires_grad]

    def update(self, parameters):
        """shadow -= (1 - d) (shadow - p) with the warm-up d = min(decay, (1+n)/(10+n))."""
        d = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            d = min(d, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            live = [p.detach() for p in self._trainable(parameters)]
            if live:
                delta = torch._foreach_sub(self.shadow_params[:len(live)], live)
                torch._foreach_mul_(delta, 1.0 - d)
                torch._foreach_sub_(self.shadow_params[:len(live)], delta)

    def copy_to(self, parameters):
        """Overwrite the given parameters with the averaged values."""
        with torch.no_grad():
            for avg, p in zip(self.shadow_params, self._trainable(parameters)):
                p.data.copy_(avg.data)

    def store(self, parameters):
        """Remember the current values so `restore` can put them back."""
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):
        with torch.no_grad():
            for saved, p in zip(self.collected_params, parameters):
                p.data.copy_(saved.data)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict['decay']
        self.num_updates = state_dict['num_updates']
        self.shadow_params = state_dict['shadow_params']
