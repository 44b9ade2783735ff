"""Host-side mirror of the reference task-model interface (src/modeling/alpro_models.py):

    AlproForVideoTextRetrieval(config, video_enc_cfg, input_format='RGB')      :727-914
    AlproForPretrain(config, video_enc_cfg, input_format='RGB')                :58-387
    Prompter(config, video_enc_cfg, input_format='RGB')                        :389-630

Same constructor arguments, batch keys, output dict keys, state_dict names/shapes, `forward_inference`,
`build_text_prompts`, `load_separate_ckpt`. `config` may be a transformers BertConfig or any object / dict exposing the
keys of config_release/base_model.json. Compute runs on the CUDA kernels through AlproEngine; there is no CPU path:
calling forward on a CPU module raises.
"""
import logging
import weakref

import torch
from torch import nn

from . import synth
from .engine import AlproEngine

LOGGER = logging.getLogger("alpro_b200")
_LIVE = weakref.WeakSet()


def live_models():
    """Every AlproBaseModel instance alive in this process (the horovod stand-in's DistributedOptimizer uses it to
    route gradient averaging through the model's flat gradient store)."""
    return list(_LIVE)


# ---------------------------------------------------------------------------------------------------------------------
# checkpoint helpers (reference: src/modeling/timesformer/helpers.py:26-56, 315-375; src/utils/load_save.py:73-140)
# ---------------------------------------------------------------------------------------------------------------------
def _nearest_index(n_out, n_in):
    """Source index of F.interpolate(mode='nearest'): floor(dst * n_in / n_out)."""
    return (torch.arange(n_out, dtype=torch.float32) * (n_in / n_out)).floor().long().clamp_(max=n_in - 1)


def resize_spatial_embedding(pos_embed, num_patches):
    """helpers.py:355-367: the cls slot is kept, the patch slots are nearest-resized as ONE flattened sequence."""
    pos_embed = pos_embed.detach()
    idx = _nearest_index(num_patches, pos_embed.shape[1] - 1) + 1
    return torch.cat([pos_embed[:, :1], pos_embed[:, idx]], dim=1)


def resize_temporal_embedding(time_embed, num_frames):
    """helpers.py:370-375."""
    time_embed = time_embed.detach()
    return time_embed[:, _nearest_index(num_frames, time_embed.shape[1])]


def read_timesformer_checkpoint(path):
    """helpers.load_state_dict (helpers.py:26-56): accepts {'state_dict': ...} ('module.' prefix stripped),
    {'model_state': ...} ('model.' prefix stripped) or a bare state dict."""
    ckpt = torch.load(path, map_location="cpu")
    if isinstance(ckpt, dict) and "state_dict" in ckpt:
        return {(k[7:] if k.startswith("module") else k): v for k, v in ckpt["state_dict"].items()}
    if isinstance(ckpt, dict) and "model_state" in ckpt:
        return {(k[6:] if k.startswith("model") else k): v for k, v in ckpt["model_state"].items()}
    return dict(ckpt)


def _report_keys(what, loaded, own):
    missing = sorted(k for k in own if k not in loaded)
    unexpected = sorted(k for k in loaded if k not in own)
    if missing or unexpected:
        LOGGER.warning("%s: %d keys of the model not in the checkpoint %s; %d checkpoint keys not in the model %s", what,
                       len(missing), missing[:8], len(unexpected), unexpected[:8])
    return missing, unexpected


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
    """Plain container used to reproduce the reference's module tree (and therefore its state_dict keys). Containers
    with numeric children (`blocks`, `encoder.layer`, `mpm_head`) index like the reference's ModuleList / Sequential."""

    def __getitem__(self, i):
        try:
            return self._modules[str(int(i))]
        except (KeyError, ValueError, TypeError):
            raise IndexError(i)

    def __len__(self):
        return sum(1 for k in self._modules if k.isdigit())

    def __iter__(self):
        return iter(self._modules[k] for k in sorted((k for k in self._modules if k.isdigit()), key=int))


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
        _LIVE.add(self)

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
        self.engine.grad_alloc = getattr(old, "grad_alloc", None)      # data-parallel hooks survive the switch
        self.engine.neg_seed = old.neg_seed
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

    def _auto_attach(self):
        """Under an initialised multi-rank process group (torchrun, or hvd.init() of the horovod stand-in) the VTC
        feature exchange and the overlapped gradient averaging attach by themselves, as the reference model picks up
        Horovod implicitly (alpro_models.py:110-111)."""
        import torch.distributed as dist
        from .engine import LocalComm
        if isinstance(self.engine.comm, LocalComm) and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size() > 1:
            from . import comm
            comm.attach(self)

    def _run(self, batch):
        self._check_device(batch)
        self._auto_attach()
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

    def load_visual_weights(self, path):
        """TimeSformer.load_state_dict(path) -> load_pretrained_kinetics (vit.py:514-533, helpers.py:315-352): keys
        relative to `visual_encoder.model`, the 400-way classifier of the checkpoint is ignored, pos/time embeddings
        are nearest-resized AT LOAD TIME to this model's grid / frame count; strict."""
        sd = read_timesformer_checkpoint(path)
        vm = self.visual_encoder.model
        own = vm.state_dict()
        sd["head.weight"], sd["head.bias"] = own["head.weight"], own["head.bias"]
        n_patches = own["pos_embed"].shape[1] - 1
        if "pos_embed" in sd and sd["pos_embed"].shape[1] != n_patches + 1:
            sd["pos_embed"] = resize_spatial_embedding(sd["pos_embed"], n_patches)
        if "time_embed" in sd and sd["time_embed"].shape[1] != own["time_embed"].shape[1]:
            sd["time_embed"] = resize_temporal_embedding(sd["time_embed"], own["time_embed"].shape[1])
        missing, unexpected = _report_keys("visual weights", sd, own)
        bad = [k for k in own if k in sd and tuple(sd[k].shape) != tuple(own[k].shape)]
        if missing or unexpected or bad:
            # the reference logs 'Error in loading Kinetics pre-trained weights' and carries on with random weights
            # (helpers.py:347-352); a silent random encoder is never what the caller wants, so this raises
            raise RuntimeError(f"visual checkpoint {path!r} does not match the TimeSformer: missing {missing[:5]}, "
                               f"unexpected {unexpected[:5]}, shape mismatch {bad[:5]}")
        vm.load_state_dict(sd, strict=True)

    def load_separate_ckpt(self, visual_weights_path=None, bert_weights_path=None):
        """alpro_models.py:45-51 (bert_weights_path is ignored there too: BERT comes from from_pretrained)."""
        if visual_weights_path:
            self.load_visual_weights(visual_weights_path)

    # ---- feature-level entry points of the reference classes
    def _forward_visual_embeds(self, visual_inputs):
        """alpro_models.py:186-194: [B,T,3,H,W] -> video_embeds [B, 1+N, d] (no gradient: the hand-written backward is
        reached through forward(batch) only)."""
        self._check_device({"visual_inputs": visual_inputs})
        with torch.no_grad():
            return self.engine.visual_features(self._tensor_dict(), visual_inputs)

    def _forward_text_feats(self, batch):
        """alpro_models.py:196-207: text_embeds [B,L,h], normalised text_feat [B,256] (no gradient)."""
        with torch.no_grad():
            return self.engine.text_features(self._tensor_dict(), batch["text_input_ids"], batch["text_input_mask"])


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
    """Teacher that turns a clip into soft entity labels (alpro_models.py:389-630)."""
    kind = "prompter"

    def __init__(self, config, video_enc_cfg, input_format="RGB"):
        super().__init__(config, input_format=input_format, video_enc_cfg=video_enc_cfg)
        self.entity_num = self._cfg["num_entities"]
        self.prompt_initialized = False
        self.ignore_threshold = 0.2         # compared with the argmax INDEX in the reference (:527), kept as is

    def load_pretrained_weights_without_prompts(self, ckpt_path):
        """alpro_models.py:404-428: everything but the *_prompt_feat buffers, non-strict, differences logged."""
        LOGGER.info("Loading weights for teacher model.")
        loaded = torch.load(ckpt_path, map_location="cpu")
        _report_keys("teacher weights", loaded, self.state_dict())
        self.load_state_dict({k: v for k, v in loaded.items() if "prompt_feat" not in k}, strict=False)

    def build_text_prompts(self, prompts):
        """alpro_models.py:430-507: encode every prompt (chunks of 10 000) with the text encoder, project + normalise the
        [CLS] output, average over the templates of each entity -> video_prompt_feat / image_prompt_feat [E,256]."""
        assert not self.prompt_initialized, "Repetitively building prompts?"
        if self.training:
            self.eval()
        P = self._tensor_dict()
        dev = self.temp.device
        if not self.temp.is_cuda:
            raise RuntimeError("alpro_b200 runs on CUDA only (sm_100a kernels): move the model to the device first")
        with torch.no_grad():
            for key, buf in (("batch_enc_video_prompts", "video_prompt_feat"),
                             ("batch_enc_image_prompts", "image_prompt_feat")):
                ids_all, mask_all = prompts[key].input_ids, prompts[key].attention_mask
                feats = []
                for s0 in range(0, ids_all.shape[0], 10000):
                    ids = ids_all[s0:s0 + 10000].to(dev)
                    mask = mask_all[s0:s0 + 10000].to(dev)
                    _, f = self.engine.text_features(P, ids, mask)
                    feats.append(f)
                f = torch.cat(feats, dim=0)
                n_templates = int(f.shape[0] / self.entity_num)
                f = torch.stack(f.chunk(n_templates), dim=1).mean(dim=1)
                getattr(self, buf).copy_(f)       # in place: keeps the registered buffer (and the parent's tensor map)
        self.prompt_initialized = True

    def _forward_visual_embeds(self, visual_inputs):
        """alpro_models.py:509-523: (video_embeds, normalised video_feat)."""
        self._check_device({"visual_inputs": visual_inputs})
        with torch.no_grad():
            return self.engine.visual_feat(self._tensor_dict(), visual_inputs)

    def get_pseudo_labels(self, batch):
        """alpro_models.py:531-551: (soft labels [B,E], ignore mask bool [B]); always eval + no_grad."""
        if self.training:
            self.eval()
        self._check_device({"visual_inputs": batch["crop_visual_inputs"]})
        with torch.no_grad():
            return self.engine.pseudo_labels(self._tensor_dict(), batch["crop_visual_inputs"], batch.get("type", "video"))

    def forward_feats(self, batch):
        """alpro_models.py:597-630 (no gradient; training the teacher goes through forward(batch))."""
        self._check_device(batch)
        with torch.no_grad():
            P = self._tensor_dict()
            self.engine.clamp_temp(P)
            ve, vf = self.engine.visual_feat(P, batch["visual_inputs"])
            te, tf = self.engine.text_features(P, batch["text_input_ids"], batch["text_input_mask"])
        return ve, vf, te, tf

    def forward(self, batch):
        """alpro_models.py:553-595: the video-text contrastive step that trains the teacher
        (run_pretrain_contrastive_only.py): itc_loss (autograd-connected), itc_labels, i2t_scores, t2i_scores."""
        out = self._run(batch)
        return _public(out, ("itc_loss", "itc_labels", "i2t_scores", "t2i_scores"))


class AlproForPretrain(AlproBaseModel):
    kind = "pretrain"

    def __init__(self, config, video_enc_cfg, input_format="RGB"):
        self._ctor = (config, video_enc_cfg, input_format)
        super().__init__(config, input_format=input_format, video_enc_cfg=video_enc_cfg)
        self.use_mask_prob = 0
        for p in self.prompter.parameters():      # teacher is frozen (eval + no_grad in the reference, :532-535)
            p.requires_grad_(False)

    def _spec(self):
        # the `prompter.` subtree is a real Prompter module (same state_dict keys)
        full = synth.model_spec(self.kind, self._cfg, self._vis, self._cfg.get("num_entities"))
        return type(full)((k, v) for k, v in full.items() if not k.startswith("prompter."))

    def _make(self, temp):
        config, video_enc_cfg, input_format = self._ctor
        super()._make(temp)
        self.prompter = Prompter(config, video_enc_cfg, input_format)      # alpro_models.py:63

    def forward(self, batch):
        out = self._run(batch)
        res = _public(out, ("itc_loss", "mlm_scores", "mlm_loss", "mlm_labels", "itm_scores", "itm_loss", "itm_labels",
                            "mpm_loss", "mpm_logits", "mpm_labels"))
        res.update({k: v for k, v in out.items() if k.startswith("_")})
        return res

    def build_text_prompts(self, prompts):
        """alpro_models.py:73-74."""
        self.prompter.build_text_prompts(prompts)

    def get_pseudo_labels(self, batch):
        """alpro_models.py:76-77."""
        return self.prompter.get_pseudo_labels(batch)

    def load_separate_ckpt(self, visual_weights_path=None, bert_weights_path=None, prompter_weights_path=None):
        """alpro_models.py:375-387."""
        if visual_weights_path:
            self.load_visual_weights(visual_weights_path)
        if prompter_weights_path is not None:
            self.prompter.load_pretrained_weights_without_prompts(prompter_weights_path)
