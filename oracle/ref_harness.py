"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (/root/reference, read-only) so that golden fixtures
can be generated from it and the oracle restatement (oracle/alpro_oracle.py) can be pinned against it.

Only usable in the build container (where /root/reference exists); nothing on the GPU box may import this module.
The reference pins transformers==4.11.3 / numpy 1.x and needs apex, horovod, ujson, tensorboardX — all absent here —
so a small set of compatibility shims is installed *before* importing src.modeling.* (SURVEY.md §8c). No reference
source is copied or edited.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("ALPRO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "modeling"))


class _HvdState:
    rank = 0
    size = 1
    allgather = None  # optional callable(tensor) -> tensor, for W>1 emulation through torch.distributed


def _install_shims():
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    import transformers.file_utils as fu

    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # alpro_models.py:293-294,824-825 (numpy 2 removed the alias)
    for name in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())

    def _noop_decorator(*a, **k):
        def deco(fn):
            return fn
        return deco

    for name in ("add_code_sample_docstrings", "add_start_docstrings", "add_start_docstrings_to_model_forward",
                 "replace_return_docstrings"):
        setattr(fu, name, _noop_decorator)
    if not hasattr(fu, "ModelOutput"):
        from transformers.utils import ModelOutput
        fu.ModelOutput = ModelOutput

    # fake third-party modules that are imported but not exercised by the modelling path
    def _mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    if "apex" not in sys.modules:
        _mod("apex")
        _mod("apex.normalization")
        _mod("apex.normalization.fused_layer_norm", FusedLayerNorm=torch.nn.LayerNorm)
        _mod("apex.amp")
    if "horovod" not in sys.modules:
        def allgather(t):
            if _HvdState.allgather is not None:
                return _HvdState.allgather(t)
            return t
        hvd = _mod("horovod.torch", allgather=allgather, local_rank=lambda: _HvdState.rank,
                   rank=lambda: _HvdState.rank, size=lambda: _HvdState.size, init=lambda: None)
        h = _mod("horovod")
        h.torch = hvd
    if "ujson" not in sys.modules:
        sys.modules["ujson"] = json
    if "tensorboardX" not in sys.modules:
        class SummaryWriter:  # noqa
            def __init__(self, *a, **k):
                pass
        _mod("tensorboardX", SummaryWriter=SummaryWriter)
    for missing in ("easydict", "decord", "av", "lmdb"):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:
                _mod(missing)


_loaded = {}


def load_reference():
    """Returns a namespace with the reference classes (alpro_models, vit, xbert modules)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import src.modeling.xbert as xbert

    P = xbert.BertPreTrainedModel
    # transformers 5.x: init_weights()/post_init plumbing changed; reference calls self.init_weights() (xbert.py:852,1354)
    def init_weights(self):
        if getattr(self, "_alpro_in_init", False):
            return
        self._alpro_in_init = True
        try:
            self.post_init()
        finally:
            self._alpro_in_init = False
    P.init_weights = init_weights
    P.get_head_mask = lambda self, hm, n, *a, **k: [None] * n  # removed in v5; xbert.py:1042
    xbert.BertForMaskedLM.from_pretrained = classmethod(lambda cls, name, config=None, **kw: cls(config))

    import src.modeling.alpro_models as am
    import src.modeling.timesformer.vit as vit
    _loaded.update(am=am, vit=vit, xbert=xbert, hvd=_HvdState)
    return _loaded


def tie_mlm_decoder(model):
    """transformers 4.11.3 ties cls.predictions.decoder.weight to the word embeddings (xbert.py:1354-1360); 5.x does
    not do it for this fork, so the harness ties explicitly."""
    te = model.text_encoder
    te.cls.predictions.decoder.weight = te.bert.embeddings.word_embeddings.weight
    if hasattr(model, "prompter"):
        tie_mlm_decoder(model.prompter)


class TinyTimeSformer(torch.nn.Module):
    """The reference hard-codes embed_dim=768/depth=12 inside TimeSformer (vit.py:445-462). For the small parity
    configuration we instantiate the reference's own VisionTransformer with other dims and reuse the reference's
    TimeSformer.forward_features unbound (vit.py:475-503) so that no reference logic is restated here."""

    def __init__(self, vit_mod, model_cfg, embed_dim, depth, num_heads, **_):
        super().__init__()
        from functools import partial
        self.img_size = model_cfg["img_size"]
        self.patch_size = model_cfg["patch_size"]
        self.num_frames = model_cfg["num_frm"]
        self.model = vit_mod.VisionTransformer(
            img_size=self.img_size, num_classes=model_cfg.get("num_classes", 400), patch_size=self.patch_size,
            embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4, qkv_bias=True,
            norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=model_cfg["drop_rate"],
            attn_drop_rate=model_cfg["attn_drop_rate"], drop_path_rate=model_cfg["drop_path_rate"],
            num_frames=self.num_frames, attention_type="divided_space_time")
        self._ff = vit_mod.TimeSformer.forward_features

    def forward_features(self, x, return_all_tokens=True, pooling="temporal"):
        return self._ff(self, x, return_all_tokens=return_all_tokens, pooling=pooling)


def build_reference_model(kind, bert_cfg_dict, video_cfg, vis_dims=None, num_entities=None):
    """kind: 'retrieval' | 'pretrain' | 'prompter'. vis_dims=(embed_dim, depth, heads) overrides the hard-coded
    768/12/12 (requires bert hidden_size == embed_dim, since fusion concatenates text and video tokens)."""
    R = load_reference()
    am, vit, xbert = R["am"], R["vit"], R["xbert"]
    cfg = xbert.BertConfig(**bert_cfg_dict)
    if num_entities is not None:
        cfg.num_entities = num_entities
    cls = {"retrieval": am.AlproForVideoTextRetrieval, "pretrain": am.AlproForPretrain, "prompter": am.Prompter}[kind]
    video_cfg = dict(video_cfg)
    video_cfg.setdefault("cls", "TimeSformer")
    if vis_dims is None:
        model = cls(cfg, video_enc_cfg=video_cfg)
    else:
        d, depth, heads = vis_dims
        assert d == cfg.hidden_size
        orig = am.TimeSformer
        am.TimeSformer = lambda model_cfg, input_format="RGB", cross_attention_config=None, **kw: TinyTimeSformer(
            vit, model_cfg, d, depth, heads)
        orig_linear = torch.nn.Linear
        try:
            # vision_width is hard-coded to 768 (alpro_models.py:33-34): swap vision_proj after construction
            model = cls(cfg, video_enc_cfg=video_cfg)
        finally:
            am.TimeSformer = orig
        def fix(m):
            m.vision_proj = orig_linear(d, 256)
            if hasattr(m, "prompter"):
                fix(m.prompter)
        fix(model)
    tie_mlm_decoder(model)
    return model


class FixedNegatives:
    """Context manager replacing torch.multinomial by a deterministic rule (argmax of the weights), recording the
    indices drawn, so that both sides of a parity test use identical hard negatives (alpro_models.py:301-316,833-844)."""

    def __init__(self):
        self.drawn = []

    def __enter__(self):
        self._orig = torch.multinomial
        def fake(w, n, *a, **k):
            idx = torch.argmax(w, dim=-1, keepdim=True)
            self.drawn.append(int(idx.item()))
            return idx
        torch.multinomial = fake
        return self

    def __exit__(self, *exc):
        torch.multinomial = self._orig
        return False
