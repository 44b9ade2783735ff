"""`apex.amp` stand-in: the reference trainers call amp.initialize(model, optimizer, enabled=cfg.fp16, opt_level='O2')
with fp16 = 0 in every released config, i.e. as a pass-through. alpro_b200 models manage their own
16-bit GEMM operands and loss scale (AlproEngine), so the stand-in is the identity the trainers already run with:

    initialize(model, optimizer=None, enabled=..., opt_level=..., **kw)   run_video_retrieval.py:329-331, 784-786
    scale_loss(loss, optimizer, delay_unscale=False)  (context manager)   run_video_retrieval.py:439-444
    master_params(optimizer)                                              run_video_retrieval.py:473-476
    state_dict() / load_state_dict(sd)                                    src/utils/load_save.py:262,276,331,346
"""
import contextlib


def initialize(models, optimizers=None, enabled=True, opt_level="O1", **kwargs):
    if optimizers is None:
        return models
    return models, optimizers


@contextlib.contextmanager
def scale_loss(loss, optimizers, loss_id=0, model=None, delay_unscale=False, delay_overflow_check=False):
    yield loss


def master_params(optimizer):
    for group in optimizer.param_groups:
        for p in group["params"]:
            yield p


def state_dict(destination=None):
    return {} if destination is None else destination


def load_state_dict(state_dict):
    return None
