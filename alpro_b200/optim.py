"""Fused AdamW + global-norm clipping on the flat gradient buffer.

Replaces `clip_grad_norm_(...)` + `AdamW.step()` of the reference trainers (run_video_retrieval.py:473-490,
src/optimization/adamw.py:40-103) with two kernel launches over flat fp32 buffers. The parameters of the model are
re-pointed to views of one flat buffer laid out exactly like the engine's GradStore, so `grad`, `exp_avg`, `exp_avg_sq`
and the parameters are element-aligned.

Parameters that receive no gradient on a path (`text_encoder.cls.*` and the 400-way `visual_encoder.model.head` in
retrieval) sit in the store with zero gradients and take the decoupled weight decay — exactly what the reference
trainer does to them: it calls `zero_none_grad(model)` before the optimizer step (run_video_retrieval.py:443,
src/utils/misc.py:28-31), so their `p.grad` is a zero tensor, not None, when AdamW reaches `:96-98`."""
import torch

from . import ops
from .engine import GradStore, bert_grad_groups


class FusedAdamW:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 max_grad_norm=-1.0):
        self.model = model
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.correct_bias, self.max_grad_norm = correct_bias, max_grad_norm
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and not n.startswith("prompter.")]
        dev = named[0][1].device
        layout = GradStore(named, dev, bert_grad_groups("text_encoder.", model.engine.cfg))
        self.offsets = layout.offsets
        self.numel = layout.flat.numel()
        assert self.numel % 4 == 0
        self.flat = layout.flat          # reuse the zero-filled buffer as parameter storage
        with torch.no_grad():
            for n, p in named:
                o, cnt, shape = self.offsets[n]
                view = self.flat[o:o + cnt].view(shape)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        # number of updates actually applied, kept on the device: a step the kernel skips (non-finite gradient norm
        # after an fp16 overflow) does not advance the bias-correction schedule
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step_size_dev = torch.zeros(1, device=dev, dtype=torch.float32)

    def step(self):
        G = self.model.engine.last_grads
        if G is None:
            raise RuntimeError("FusedAdamW.step(): run loss.backward() first")
        assert G.offsets == self.offsets, "gradient layout changed"
        b1, b2 = self.betas
        # the squared gradient norm is always computed: it drives clipping and makes the kernel skip the update when the
        # fp16 backward overflowed (non-finite norm)
        self.gnorm_sq.zero_()
        ops.sumsq(G.flat, self.gnorm_sq)
        clip = self.max_grad_norm if (self.max_grad_norm is not None and self.max_grad_norm > 0) else -1.0
        ops.adamw_prepare(self.gnorm_sq, self.lr, b1, b2, self.correct_bias, self.step_dev, self.step_size_dev)
        ops.adamw_step_dev(self.flat, G.flat, self.exp_avg, self.exp_avg_sq, b1, b2, self.eps, self.step_size_dev,
                           self.lr * self.wd, self.gnorm_sq, clip)
        self.model.engine.W.invalidate()     # 16-bit operand copies are stale now

    def update_loss_scale(self, growth_interval=1000):
        """Dynamic loss scaling for the fp16 backward (host sync: call every few steps, not every step). Halves the
        engine's scale after an overflowed step (which the kernel skipped), doubles it after `growth_interval` good ones."""
        eng = self.model.engine
        if not bool(torch.isfinite(self.gnorm_sq).item()):
            eng.S = max(eng.S * 0.5, 1.0)
            self._good = 0
        else:
            self._good = getattr(self, "_good", 0) + 1
            if self._good >= growth_interval and eng.S < 65536.0:
                eng.S *= 2.0
                self._good = 0
        return eng.S

    @property
    def step_count(self):
        """Updates applied so far (host sync)."""
        return int(self.step_dev.item())

    def grad_norm(self):
        """sqrt of the last computed squared gradient norm (device tensor; no sync)."""
        return self.gnorm_sq.sqrt()

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None
