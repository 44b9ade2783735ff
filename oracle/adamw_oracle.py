"""ORACLE — TEST INFRASTRUCTURE. CPU restatement of the reference optimizer step: torch.nn.utils.clip_grad_norm_ as used
at src/tasks/run_video_retrieval.py:473-476 followed by AdamW.step (src/optimization/adamw.py:40-103). Pinned against
the reference class itself in tests/test_optim.py::test_adamw_oracle_matches_reference (build container only)."""
import math

import torch


def clip_coef(grads, max_norm):
    """clip_grad_norm_: total L2 norm over all grads; coef = max_norm / (norm + 1e-6), applied only when < 1."""
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
    c = max_norm / (total + 1e-6)
    return (c if c < 1.0 else 1.0), total


def adamw_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
    """adamw.py:73-98 on one tensor (in place on p, m, v)."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
