"""Host-side mirror of the reference task-model interface (src/modeling/alpro_models.py):

    AlproForVideoTextRetrieval(config, video_enc_cfg, input_format='RGB')      :727-914
    AlproForPretrain(config, video_enc_cfg, input_format='RGB')                :58-387
    Prompter(config, video_enc_cfg, input_format='RGB')                        :389-630

Same constructor arguments, batch keys, output dict keys, state_dict names/shapes, `forward_inference`,
`build_text_prompts`, `load_separate_ckpt`. `config` may be a transformers BertConfig or any object / dict exposing the
keys of config_release/base_model.json. Compute runs on the CUDA kernels through AlproEngine; there is no CPU path:
calling forward on a CPU module raises.
"""
import torch
from torch import nn

from . import synth
from .engine import AlproEngine


def _cfg_dict(config):
    if isinstance(config, dict):
        return dict(config)
    if hasattr(config, "to_dict"):
        d = dict(config.to_dict())
    else:
        d = {k: getattr(config, k) for k in dir(config) if not k.startswith("_") and not callable(getattr(config, k))}
    for k in ("fusion_layer", "encoder_width", "itc_token_type", "num_entities"):
        if hasattr(config, k):
            d[k] = getattr(config, k)
    return d


def _vis_from_cfg(video_enc_cfg, d=768, depth=12, heads=12):
    """TimeSformer dims are hard-coded in the reference (vit.py:445-462); optional keys allow the small parity
    configuration (embed_dim/depth/num_heads)."""
    vis = dict(d=video_enc_cfg.get("embed_dim", d), depth=video_enc_cfg.get("depth", depth),
               heads=video_enc_cfg.get("num_heads", heads), T=video_enc_cfg["num_frm"],
               img=video_enc_cfg["img_size"], patch=video_enc_cfg["patch_size"],
               drop_path_rate=float(video_enc_cfg.get("drop_path_rate", 0.0)))
    # ImageNorm constants of the uint8 input path (config img_pixel_mean / img_pixel_std, e.g.
    # config_release/msrvtt_ret.json:19-20; the trainer passes them to ImageNorm, src/datasets/data_utils.py:437-457)
    if "img_pixel_mean" in video_enc_cfg:
        vis["img_mean"] = tuple(float(v) for v in video_enc_cfg["img_pixel_mean"])
    if "img_pixel_std" in video_enc_cfg:
        vis["img_std"] = tuple(float(v) for v in video_enc_cfg["img_pixel_std"])
    return vis


def _default_dtype():
    """ALPRO_DTYPE=bf16|fp16 selects the operand format of newly built models (default fp16)."""
    import os
    v = os.environ.get("ALPRO_DTYPE", "fp16").lower()
    if v in ("bf16", "bfloat16"):
        return torch.bfloat16
    if v in ("fp16", "float16", "half"):
        return torch.float16
    raise ValueError(f"ALPRO_DTYPE={v!r}: expected fp16 or bf16")


class _Holder(nn.Module):
    """Plain container used to reproduce the reference's module tree (and therefore its state_dict keys)."""


def _build_tree(root, spec, buffers=()):
    params = {}
    for name, shape in spec.items():
        canon = synth.canonical_name(name)
        parts = name.split(".")
        mod = root
        for part in parts[:-1]:
            if not hasattr(mod, part):
                mod.add_module(part, _Holder())
            mod = getattr(mod, part)
        leaf = parts[-1]
        if any(name.endswith(b) for b in buffers):
            if name.endswith("position_ids"):
                val = torch.arange(shape[1]).expand((1, -1)).clone()
            else:
                val = torch.rand(shape)
            mod.register_buffer(leaf, val)
            continue
        if canon in params:                      # tied parameter: register the same object under the alias
            mod.register_parameter(leaf, params[canon])
            continue
        p = nn.Parameter(torch.zeros(shape))
        mod.register_parameter(leaf, p)
        params[name] = p
    return params


def _init_like_reference(model):
    """Random init in the spirit of the reference (trunc_normal .02 / BERT normal .02, LN = 1/0, temporal_fc = 0 for
    blocks > 0; vit.py:285-307, xbert.py:_init_weights). Real use loads a checkpoint through load_state_dict."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            leaf = n.split(".")[-1]
            parent = n.split(".")[-2].lower() if "." in n else ""
            if n.endswith("temp"):
                p.fill_(0.07)
            elif "norm" in parent:
                p.fill_(1.0 if leaf == "weight" else 0.0)
            elif leaf == "bias":
                p.zero_()
            else:
                nn.init.trunc_normal_(p, std=0.02)
            if ".temporal_fc." in n and ".blocks.0." not in n:
                p.zero_()


class _StepFn(torch.autograd.Function):
    """Bridges the hand-written forward/backward into autograd so that the reference trainers' `loss.backward()`
    (run_video_retrieval.py:432-442) populates `p.grad` of every nn.Parameter."""

    @staticmethod
    def forward(ctx, model, batch, names, need, *params):
        P = model._tensor_dict()
        out, ectx = model.engine.forward(P, batch, need_grad=need, training=model.training)
        ctx.model, ctx.ectx, ctx.names = model, ectx, names
        ctx.loss_keys = [k for k in ("itc_loss", "itm_loss", "mlm_loss", "mpm_loss") if out.get(k) is not None]
        model._last_out = out
        losses = tuple(out[k] for k in ctx.loss_keys)
        return losses

    @staticmethod
    def backward(ctx, *grads):
        model = ctx.model
        if ctx.ectx is None:
            raise RuntimeError("backward called on a forward that ran without gradient tracking")
        g = {k: (gi.contiguous().float() if gi is not None else None) for k, gi in zip(ctx.loss_keys, grads)}
        P = model._tensor_dict()
        c = model._param_cache()
        named = [(n, p) for n, p in zip(c["names"], c["params"]) if p.requires_grad]
        G = model.engine.backward(P, ctx.ectx, named, g)
        ctx.ectx = None
        return (None, None, None, None) + model._publish_grads(G, ctx.names, named)


class AlproBaseModel(nn.Module):
    kind = "retrieval"

    def __init__(self, config=None, input_format="RGB", video_enc_cfg=None, temp=0.07):
        super().__init__()
        assert input_format == "RGB", "Official TimeSformer uses RGB input."       # vit.py:441
        self.bert_config = config
        self._cfg = _cfg_dict(config)
        self._vis = _vis_from_cfg(video_enc_cfg)
        self.itc_token_type = self._cfg.get("itc_token_type", "cls")
        assert self.itc_token_type == "cls", "Support CLS tokens for ITC only"      # alpro_models.py:113
        self._make(temp)

    def _spec(self):
        return synth.model_spec(self.kind, self._cfg, self._vis, self._cfg.get("num_entities"))

    def _make(self, temp):
        spec = self._spec()
        _build_tree(self, spec, buffers=("position_ids", "prompt_feat"))
        _init_like_reference(self)
        with torch.no_grad():
            self.temp.fill_(temp)
        self.engine = AlproEngine(self.kind, self._cfg, self._vis, dtype=_default_dtype(),
                                  num_entities=self._cfg.get("num_entities"))
        self._last_out = None

    def set_compute_dtype(self, dtype, loss_scale=4096.0):
        """GEMM operand / saved-activation format: torch.float16 (default; backward carries a static loss scale) or
        torch.bfloat16 (no loss scale). Accumulation, residual stream and statistics are fp32 in both."""
        if dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("compute dtype must be torch.float16 or torch.bfloat16")
        old = self.engine
        self.engine = AlproEngine(self.kind, self._cfg, self._vis, dtype=dtype, loss_scale=loss_scale,
                                  num_entities=self._cfg.get("num_entities"))
        self.engine.sampler, self.engine.comm = old.sampler, old.comm
        self.engine.grad_ready_hook = old.grad_ready_hook
        return self

    # ---- plumbing
    # Name -> tensor maps are built once and reused every step (walking named_parameters()/state_dict() cost ~19 ms of
    # host time per training step, tools/host_profile.py). Parameters keep their identity under .to()/.half()/
    # load_state_dict (in-place); buffers are re-created by Module._apply, which therefore drops the cache.
    def _param_cache(self):
        c = self.__dict__.get("_pcache")
        if c is None:
            d = {n: p for n, p in self.named_parameters()}
            d.update({n: b for n, b in self.named_buffers()})
            for n, t in list(self.state_dict(keep_vars=True).items()):   # aliases of tied parameters
                if n not in d:
                    d[n] = t
            names, params = [], []
            for n, p in self.named_parameters():
                if not n.startswith("prompter."):
                    names.append(n)
                    params.append(p)
            c = {"tensors": d, "names": names, "params": params}
            self.__dict__["_pcache"] = c
        return c

    def invalidate_cache(self):
        """Call after replacing a parameter/buffer OBJECT by hand (in-place updates never need it)."""
        self.__dict__["_pcache"] = None

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self.invalidate_cache()
        return out

    def _tensor_dict(self):
        return self._param_cache()["tensors"]

    def set_image_norm(self, mean, std):
        """ImageNorm constants (img_pixel_mean / img_pixel_std of the task config) for raw uint8 `visual_inputs`."""
        self._vis["img_mean"] = tuple(float(v) for v in mean)
        self._vis["img_std"] = tuple(float(v) for v in std)
        for enc in (self.engine.visual, getattr(self.engine, "t_visual", None)):
            if enc is not None:
                enc.img_mean, enc.img_std = self._vis["img_mean"], self._vis["img_std"]
        return self

    def invalidate_operands(self):
        """Drop the 16-bit operand copies (needed only after an out-of-band `.data` update between two no-grad
        forwards; training steps refresh them by themselves)."""
        self.engine.W.invalidate()

    def _publish_grads(self, G, names, named):
        """Hand the flat gradient store of one backward pass to autograd.

        First backward after `p.grad = None` (the reference trainers' zero_none_grad / set_to_none): the returned views
        become `p.grad`, so `p.grad` ALIASES `engine.last_grads.flat` — the buffer comm.allreduce_gradients and
        FusedAdamW work on. Gradient accumulation (`gradient_accumulation_steps` > 1, or `zero_grad(set_to_none=False)`):
        `p.grad` still aliases the store of the first micro-step, so the new store is added into it with ONE flat add
        (after its own overlapped all-reduce, if any, has been joined: averaging is linear) and that accumulated store
        stays `engine.last_grads`. A `p.grad` that does not alias our store (assigned by hand) falls back to autograd's
        per-tensor accumulation and is flagged, so that allreduce_gradients reduces the `p.grad` tensors themselves."""
        eng = self.engine
        acc = getattr(eng, "grad_accum", None)
        live = [(n, p) for n, p in named if p.grad is not None]
        if not live:
            eng.grad_accum = eng.last_grads = G
            self._grads_aliased = True
            return tuple(G[n] if n in G else None for n in names)
        aliased = (acc is not None and acc.offsets == G.offsets and acc.flat.device == G.flat.device
                   and len(live) == len(named)
                   and all(p.grad.data_ptr() == acc[n].data_ptr() for n, p in named))
        if aliased:
            red = getattr(self, "_grad_reducer", None)
            if red is not None and red._G is G:
                red.finish()
            acc.flat.add_(G.flat)
            eng.last_grads = acc
            self._grads_aliased = True
            return tuple(None for _ in names)
        self._grads_aliased = False
        eng.grad_accum = None
        return tuple(G[n] if n in G else None for n in names)

    def _check_device(self, batch):
        v = batch["visual_inputs"]
        if not v.is_cuda or not self.temp.is_cuda:
            raise RuntimeError("alpro_b200 runs on CUDA only (sm_100a kernels); move the model and batch to a B200 "
                               "device — there is no CPU fallback.")

    def _run(self, batch):
        self._check_device(batch)
        c = self._param_cache()
        names, params = c["names"], c["params"]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        losses = _StepFn.apply(self, batch, names, need, *params)
        out = dict(self._last_out)
        self._last_out = None
        keys = [k for k in ("itc_loss", "itm_loss", "mlm_loss", "mpm_loss") if out.get(k) is not None]
        for k, v in zip(keys, losses):
            out[k] = v
        return out

    def load_separate_ckpt(self, visual_weights_path=None, bert_weights_path=None):
        """alpro_models.py:45-51: loads TimeSformer weights. Accepts a state_dict file whose keys are relative to
        visual_encoder.model (pos/time embeddings are resized at run time)."""
        if visual_weights_path:
            sd = torch.load(visual_weights_path, map_location="cpu")
            sd = sd.get("model_state", sd)
            own = self.visual_encoder.model.state_dict()
            self.visual_encoder.model.load_state_dict({k: v for k, v in sd.items() if k in own and v.shape == own[k].shape},
                                                      strict=False)


def _public(out, keys):
    return {k: out.get(k) for k in keys}


class AlproForVideoTextRetrieval(AlproBaseModel):
    kind = "retrieval"

    def __init__(self, config, video_enc_cfg, input_format="RGB"):
        super().__init__(config, input_format=input_format, video_enc_cfg=video_enc_cfg)

    def forward(self, batch):
        out = self._run(batch)
        res = _public(out, ("itm_scores", "itm_loss", "itm_labels", "itc_loss"))
        res.update({k: v for k, v in out.items() if k.startswith("_")})
        return res

    def forward_inference(self, batch):
        self._check_device(batch)
        with torch.no_grad():
            return self.engine.inference(self._tensor_dict(), batch)


class Prompter(AlproBaseModel):
    kind = "prompter"

    def __init__(self, config, video_enc_cfg, input_format="RGB"):
        super().__init__(config, input_format=input_format, video_enc_cfg=video_enc_cfg)
        self.entity_num = self._cfg["num_entities"]
        self.prompt_initialized = False
        self.ignore_threshold = 0.2


class AlproForPretrain(AlproBaseModel):
    kind = "pretrain"

    def __init__(self, config, video_enc_cfg, input_format="RGB"):
        super().__init__(config, input_format=input_format, video_enc_cfg=video_enc_cfg)
        self.use_mask_prob = 0
        for p in self.prompter.parameters():      # teacher is frozen (eval + no_grad in the reference, :532-535)
            p.requires_grad_(False)

    def forward(self, batch):
        out = self._run(batch)
        res = _public(out, ("itc_loss", "mlm_scores", "mlm_loss", "mlm_labels", "itm_scores", "itm_loss", "itm_labels",
                            "mpm_loss", "mpm_logits", "mpm_labels"))
        res.update({k: v for k, v in out.items() if k.startswith("_")})
        return res

    def build_text_prompts(self, prompts):
        """Prompter.build_text_prompts (alpro_models.py:430-507): encode every prompt with the teacher's text encoder,
        normalise, average over templates -> prompter.{video,image}_prompt_feat."""
        from .engine import BertEncoder, _PrefixedCache
        P = {k[len("prompter."):]: v for k, v in self._tensor_dict().items() if k.startswith("prompter.")}
        eng = self.engine
        bert = BertEncoder("text_encoder.", eng.cfg, eng.dtype)
        cache = _PrefixedCache(eng.W, "prompter.")
        E = self._cfg["num_entities"]
        with torch.no_grad():
            for key, buf in (("batch_enc_video_prompts", "video_prompt_feat"), ("batch_enc_image_prompts", "image_prompt_feat")):
                ids = prompts[key].input_ids.cuda()
                mask = prompts[key].attention_mask.cuda().contiguous()
                feats = []
                for s in range(0, ids.shape[0], 10000):
                    i, m = ids[s:s + 10000].contiguous(), mask[s:s + 10000].contiguous()
                    x32, x16, _ = bert.embed(P, i, False)
                    te, _, _ = bert.forward(P, cache, x32, x16, eng._text_mask_add(m), i.shape[0], i.shape[1], "text", False)
                    f, _ = eng._proj_norm(P, te, i.shape[1] * eng.cfg["hidden_size"], "text_proj", i.shape[0])
                    feats.append(f)
                f = torch.cat(feats, dim=0)
                f = torch.stack(f.chunk(f.shape[0] // E), dim=1).mean(dim=1)
                getattr(self.prompter, buf).copy_(f)
        self.prompter.prompt_initialized = True
