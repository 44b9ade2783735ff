"""GPU: the teacher-side entry points of the reference (alpro_models.py:389-630) and the run-time embedding resize
(vit.py:328-355) against the CPU oracle: Prompter.forward (+ backward), forward_feats, get_pseudo_labels (stand-alone
and through AlproForPretrain), build_text_prompts, and a clip whose grid / frame count differ from the checkpoint's."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from alpro_b200 import modeling, synth  # noqa: E402
from oracle import alpro_oracle, configs  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_parity import build_cuda_model, to_cuda  # noqa: E402


def _prompter(cfg, seed):
    v = dict(cfg["video"])
    v.update(embed_dim=cfg["vis"]["d"], depth=cfg["vis"]["depth"], num_heads=cfg["vis"]["heads"])
    b = dict(cfg["bert"])
    b["num_entities"] = cfg["num_entities"]
    spec = synth.model_spec("prompter", cfg["bert"], cfg["vis"], cfg["num_entities"])
    sd = synth.synth_state_dict(spec, seed)
    m = modeling.Prompter(b, v)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def _leaves(sd):
    sd_o = {k: v.clone() for k, v in sd.items()}
    for k in list(sd_o):
        c = synth.canonical_name(k)
        if c != k:
            sd_o[k] = sd_o[c]
    for v in sd_o.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    return sd_o


def test_prompter_forward_and_backward_match_oracle():
    cfg = configs.tiny("prompter", B=3, T=2, img=64, L=8, seed=17)
    model, sd = _prompter(cfg, 17)
    batch = synth.synth_batch("pretrain", 3, 2, 64, 8, cfg["bert"]["vocab_size"], seed=17, num_entities=cfg["num_entities"])
    out = model(to_cuda(batch))
    assert set(out) == {"itc_loss", "itc_labels", "i2t_scores", "t2i_scores"}
    out["itc_loss"].backward()
    sd_o = _leaves(sd)
    ref = alpro_oracle.prompter_forward(sd_o, cfg["bert"], cfg["vis"], batch)
    ref["itc_loss"].backward()
    assert out["itc_labels"].tolist() == ref["itc_labels"].tolist()
    for k in ("itc_loss", "i2t_scores", "t2i_scores"):
        assert helpers.rel_err(out[k].detach().float().cpu(), ref[k].detach()) < 1e-3, k
    gmax = max(float(v.grad.abs().max()) for v in sd_o.values() if v.grad is not None)
    bad, errs = [], []
    for n, p in model.named_parameters():
        r = sd_o[n].grad
        if r is None or float(r.abs().max()) < 1e-6 * gmax:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-4 * gmax, n     # fusion layers / heads: untouched
            continue
        e = helpers.rel_err(p.grad.cpu(), r)
        errs.append(e)
        # the contrastive gradient reaches the encoders through the B [CLS] rows only: bias gradients are sums of B = 3
        # cancelling rows, so a few of them sit at 3-5e-2 of their (small) max where the dense losses give < 3e-2
        if e > 8e-2:
            bad.append((n, e))
    errs.sort()
    print("prompter grads: n", len(errs), "median", errs[len(errs) // 2], "max", errs[-1])
    assert len(errs) > 40 and not bad, bad[:8]
    assert errs[len(errs) // 2] < 2e-2, errs[len(errs) // 2]       # measured: median 1.05e-2, max 5.4e-2
    # forward_feats: the four feature tensors of alpro_models.py:597-630
    ve, vf, te, tf = model.forward_feats(to_cuda(batch))
    assert helpers.rel_err(ve.cpu(), ref["_video_embeds"].detach()) < 2e-3
    assert helpers.rel_err(vf.cpu(), ref["_video_feat"].detach()) < 1e-3
    assert helpers.rel_err(te.cpu(), ref["_text_embeds"].detach()) < 2e-3
    assert helpers.rel_err(tf.cpu(), ref["_text_feat"].detach()) < 1e-3


def test_pseudo_labels_stand_alone_and_through_pretrain_model():
    cfg = configs.GOLDEN["tiny_pretrain"]
    spec, sd, batch = helpers.make_inputs(cfg)
    soft_o, ign_o = alpro_oracle.pseudo_labels(sd, cfg["bert"], cfg["vis"], batch)
    model = build_cuda_model(cfg, sd)
    assert isinstance(model.prompter, modeling.Prompter)
    model.train()
    soft, ign = model.get_pseudo_labels(to_cuda(batch))            # AlproForPretrain.get_pseudo_labels :76-77
    assert not model.prompter.training                              # the teacher switches itself to eval (:532-533)
    assert ign.dtype == torch.bool and ign.cpu().tolist() == ign_o.tolist()
    assert helpers.rel_err(soft.cpu(), soft_o) < 3e-3
    for img_type in ("video", "img"):
        b = dict(batch, type=img_type)
        s2, _ = model.prompter.get_pseudo_labels(to_cuda(b))
        s_o, _ = alpro_oracle.pseudo_labels(sd, cfg["bert"], cfg["vis"], b)
        assert helpers.rel_err(s2.cpu(), s_o) < 3e-3, img_type


def test_build_text_prompts_matches_oracle():
    cfg = configs.GOLDEN["tiny_pretrain"]
    spec, sd, batch = helpers.make_inputs(cfg)
    model = build_cuda_model(cfg, sd)
    E, nt, Lp = cfg["num_entities"], 3, 6
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(4, cfg["bert"]["vocab_size"], (E * nt, Lp), generator=g)
    ids[:, 0] = 1
    mask = torch.ones(E * nt, Lp, dtype=torch.long)
    mask[::2, -2:] = 0
    ids2 = ids.flip(0).contiguous()
    prompts = {"batch_enc_video_prompts": SimpleNamespace(input_ids=ids, attention_mask=mask),
               "batch_enc_image_prompts": SimpleNamespace(input_ids=ids2, attention_mask=mask)}
    before = model.prompter.video_prompt_feat.clone()
    model.build_text_prompts(prompts)                                # AlproForPretrain.build_text_prompts :73-74
    assert model.prompter.prompt_initialized
    tsd = {k[len("prompter."):]: v for k, v in sd.items() if k.startswith("prompter.")}
    want_v = alpro_oracle.build_text_prompts(tsd, cfg["bert"], ids, mask, E)
    want_i = alpro_oracle.build_text_prompts(tsd, cfg["bert"], ids2, mask, E)
    assert not torch.equal(before, model.prompter.video_prompt_feat)
    assert helpers.rel_err(model.prompter.video_prompt_feat.cpu(), want_v) < 1e-3
    assert helpers.rel_err(model.prompter.image_prompt_feat.cpu(), want_i) < 1e-3
    with pytest.raises(AssertionError):
        model.build_text_prompts(prompts)                            # "Repetitively building prompts?" :435
    # the student step now consumes the new prompt features (same buffer objects)
    sd2 = dict(sd)
    sd2["prompter.video_prompt_feat"], sd2["prompter.image_prompt_feat"] = want_v, want_i
    _, ref = helpers.oracle_run(cfg, sd2, batch)
    out = model(to_cuda(batch))
    assert helpers.rel_err(out["mpm_labels"].cpu(), ref["mpm_labels"]) < 3e-3
    assert helpers.rel_err(out["mpm_loss"].detach().cpu(), ref["mpm_loss"].detach()) < 1e-3


@pytest.mark.parametrize("ckpt_img,ckpt_T", [(48, 4), (96, 2), (64, 1)])
def test_runtime_pos_time_resize_matches_oracle(ckpt_img, ckpt_T):
    """vit.py:328-355: a clip whose patch grid / frame count differ from the stored pos_embed / time_embed gets them
    nearest-resized inside forward_features; the gradient flows back to the stored (checkpoint-shaped) parameters."""
    cfg = configs.tiny("retrieval", B=2, T=2, img=64, L=8, seed=23)
    ck = configs.tiny("retrieval", B=2, T=ckpt_T, img=ckpt_img, L=8, seed=23)
    spec = synth.model_spec("retrieval", ck["bert"], ck["vis"], None)
    sd = synth.synth_state_dict(spec, 23)
    model = build_cuda_model(ck, sd)                                 # parameters shaped like the checkpoint
    batch = synth.synth_batch("retrieval", 2, 2, 64, 8, cfg["bert"]["vocab_size"], seed=23)   # 4x4 grid, 2 frames
    out = model(to_cuda(batch))
    (out["itc_loss"] + out["itm_loss"]).backward()
    sd_o = _leaves(sd)
    ref = alpro_oracle.retrieval_forward(sd_o, ck["bert"], ck["vis"], batch)
    (ref["itc_loss"] + ref["itm_loss"]).backward()
    for k in ("itc_loss", "itm_loss", "itm_scores"):
        assert helpers.rel_err(out[k].detach().float().cpu(), ref[k].detach()) < 1e-3, k
    assert helpers.rel_err(out["_video_embeds"].cpu(), ref["_video_embeds"].detach()) < 2e-3
    vm = model.visual_encoder.model
    for name, p in (("pos_embed", vm.pos_embed), ("time_embed", vm.time_embed)):
        r = sd_o["visual_encoder.model." + name].grad
        assert p.grad.shape == r.shape
        assert helpers.rel_err(p.grad.cpu(), r) < 3e-2, name
