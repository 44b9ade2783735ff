"""CPU: pin the oracle restatement (oracle/alpro_oracle.py) against golden vectors produced by the unmodified
reference (oracle/make_golden.py), and — when /root/reference is present — against the live reference."""
import numpy as np
import pytest
import torch

from oracle import alpro_oracle, configs
from tests import helpers

TOL = 2e-5  # fp32 CPU vs fp32 CPU (different op order only)


@pytest.mark.parametrize("name", list(configs.GOLDEN))
def test_oracle_matches_reference_golden(name):
    cfg = configs.GOLDEN[name]
    gold = helpers.load_golden(name)
    spec, sd, batch = helpers.make_inputs(cfg)
    sd_g, out = helpers.oracle_run(cfg, sd, batch, requires_grad=True)
    # hard negatives (index arithmetic): bit-exact
    drawn = list(out["_neg_video"]) + list(out["_neg_text"])
    assert drawn == gold["neg_drawn"].tolist()
    assert out["itm_labels"].tolist() == gold["out.itm_labels"].tolist()
    for k in ("itc_loss", "itm_loss", "itm_scores", "mlm_loss", "mlm_scores", "mpm_loss", "mpm_logits", "mpm_labels"):
        if "out." + k in gold:
            assert helpers.rel_err(out[k].detach(), gold["out." + k]) < TOL, k
    assert helpers.rel_err(out["_video_embeds"].detach(), gold["act.video_embeds"]) < TOL
    assert helpers.rel_err(out["_text_embeds"].detach(), gold["act.text_embeds"]) < TOL
    # gradients
    loss = sum(v for k, v in out.items() if k.endswith("_loss") and v is not None)
    loss.backward()
    gnorm = dict(zip(gold["grad_names"].tolist(), gold["grad_norms"].tolist()))
    checked = 0
    for n, ref in gnorm.items():
        g = sd_g[n].grad if n in sd_g else None
        mine = 0.0 if g is None else float(g.double().norm())
        assert abs(mine - ref) <= 1e-4 * max(ref, 1e-6) + 1e-9, (n, mine, ref)
        checked += 1
    assert checked > 50
    for k in gold:
        if k.startswith("grad."):
            n = k[5:]
            assert helpers.rel_err(sd_g[n].grad, gold[k]) < 1e-4, n


def test_token_order_bit_exact():
    """Token index = 1 + n*T + t: tag tokens through zero weights so that the embedding sum is exactly pos+time."""
    cfg = configs.GOLDEN["tiny_t4_retrieval"]
    gold = helpers.load_golden("tiny_t4_retrieval")
    spec, sd, batch = helpers.make_inputs(cfg)
    out, tokens = alpro_oracle.visual_forward(sd, "visual_encoder.model.", batch["visual_inputs"], cfg["vis"],
                                              return_tokens=True)
    head = tokens[:, : 1 + 2 * cfg["T"]]
    assert helpers.rel_err(head, gold["act.video_tokens_head"]) < TOL
    # integer-tagged embedding: with zero patch weights token (n,t) must equal pos[n+1] + time[t] exactly
    sd2 = dict(sd)
    p = "visual_encoder.model."
    sd2[p + "patch_embed.proj.weight"] = torch.zeros_like(sd[p + "patch_embed.proj.weight"])
    sd2[p + "patch_embed.proj.bias"] = torch.zeros_like(sd[p + "patch_embed.proj.bias"])
    N = (cfg["img"] // 16) ** 2
    T = cfg["T"]
    d = cfg["vis"]["d"]
    sd2[p + "pos_embed"] = (torch.arange(N + 1, dtype=torch.float32) * 100).view(1, N + 1, 1).expand(1, N + 1, d).clone()
    sd2[p + "time_embed"] = torch.arange(T, dtype=torch.float32).view(1, T, 1).expand(1, T, d).clone()
    sd2[p + "cls_token"] = torch.zeros(1, 1, d)
    x = alpro_oracle.vit_tokens(sd2, p, batch["visual_inputs"], 16)
    want = [0.0] + [100.0 * (n + 1) + t for n in range(N) for t in range(T)]
    assert x[0, :, 0].tolist() == want


def test_inference_matches_golden():
    cfg = configs.GOLDEN["tiny_retrieval"]
    gold = helpers.load_golden("tiny_retrieval")
    spec, sd, batch = helpers.make_inputs(cfg)
    b1 = {"visual_inputs": batch["visual_inputs"][:1], "text_input_ids": batch["text_input_ids"],
          "text_input_mask": batch["text_input_mask"]}
    out = alpro_oracle.inference_forward(sd, cfg["bert"], cfg["vis"], b1)
    assert helpers.rel_err(out["logits"], gold["inf.logits"]) < TOL
    assert helpers.rel_err(out["itc_scores"], gold["inf.itc_scores"]) < TOL


@pytest.mark.reference
def test_state_dict_schema_matches_reference():
    """alpro_b200.synth.model_spec must reproduce the reference model's state_dict names and shapes."""
    from alpro_b200 import synth
    from oracle import ref_harness
    for kind in ("retrieval", "pretrain"):
        cfg = configs.tiny(kind)
        model = ref_harness.build_reference_model(kind, cfg["bert"], cfg["video"],
                                                  vis_dims=(cfg["vis"]["d"], cfg["vis"]["depth"], cfg["vis"]["heads"]),
                                                  num_entities=cfg["num_entities"])
        spec = synth.model_spec(kind, cfg["bert"], cfg["vis"], cfg["num_entities"])
        ref = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert ref == {k: tuple(v) for k, v in spec.items()}


@pytest.mark.reference
def test_prompter_oracle_matches_live_reference():
    """Pins oracle.prompter_forward / pseudo_labels / build_text_prompts against the unmodified reference Prompter
    (alpro_models.py:389-630) on the tiny config: Prompter.forward outputs + gradients, get_pseudo_labels, and
    build_text_prompts (whose `.cuda()` calls are neutralised for the CPU run)."""
    from types import SimpleNamespace
    from alpro_b200 import synth
    from oracle import ref_harness
    cfg = configs.tiny("prompter", B=3, T=2, img=64, L=8, seed=17)
    model = ref_harness.build_reference_model("prompter", cfg["bert"], cfg["video"],
                                              vis_dims=(cfg["vis"]["d"], cfg["vis"]["depth"], cfg["vis"]["heads"]),
                                              num_entities=cfg["num_entities"])
    spec = synth.model_spec("prompter", cfg["bert"], cfg["vis"], cfg["num_entities"])
    sd = synth.synth_state_dict(spec, cfg["seed"])
    model.load_state_dict(sd, strict=True)
    model.eval()
    batch = synth.synth_batch("pretrain", cfg["B"], cfg["T"], cfg["img"], cfg["L"], cfg["bert"]["vocab_size"],
                              seed=cfg["seed"], num_entities=cfg["num_entities"])
    # ---- forward (contrastive objective) + gradients
    ref = model(batch)
    ref["itc_loss"].backward()
    sd_o = {k: v.clone() for k, v in sd.items()}
    for k in list(sd_o):
        c = synth.canonical_name(k)
        if c != k:
            sd_o[k] = sd_o[c]
    for v in sd_o.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    out = alpro_oracle.prompter_forward(sd_o, cfg["bert"], cfg["vis"], batch)
    out["itc_loss"].backward()
    for k in ("itc_loss", "i2t_scores", "t2i_scores"):
        assert helpers.rel_err(out[k].detach(), ref[k].detach()) < TOL, k
    assert out["itc_labels"].tolist() == ref["itc_labels"].tolist()
    checked = 0
    for n, p in model.named_parameters():
        if p.grad is None or float(p.grad.abs().max()) < 1e-9:
            continue
        assert helpers.rel_err(sd_o[n].grad, p.grad) < 1e-4, n
        checked += 1
    assert checked > 40
    # ---- pseudo labels
    soft_r, ign_r = model.get_pseudo_labels(batch)
    sd_p = {"prompter." + k: v for k, v in sd.items()}
    soft_o, ign_o = alpro_oracle.pseudo_labels(sd_p, cfg["bert"], cfg["vis"], batch)
    assert helpers.rel_err(soft_o, soft_r) < TOL and ign_o.tolist() == ign_r.tolist()
    # ---- prompts: E entities x 3 templates, template-major
    E, nt, Lp = cfg["num_entities"], 3, 6
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(4, cfg["bert"]["vocab_size"], (E * nt, Lp), generator=g)
    ids[:, 0] = 1
    mask = torch.ones(E * nt, Lp, dtype=torch.long)
    mask[::2, -2:] = 0

    class _T(torch.Tensor):
        def cuda(self, *a, **k):     # alpro_models.py:455-456 moves the chunk to the GPU; there is none here
            return self.as_subclass(torch.Tensor)

        def __getitem__(self, i):
            return torch.Tensor.__getitem__(self, i).as_subclass(_T)
    enc = SimpleNamespace(input_ids=ids.as_subclass(_T), attention_mask=mask.as_subclass(_T))
    model.prompt_initialized = False
    model.build_text_prompts({"batch_enc_video_prompts": enc, "batch_enc_image_prompts": enc})
    want = alpro_oracle.build_text_prompts(sd, cfg["bert"], ids, mask, E)
    assert helpers.rel_err(want, model.video_prompt_feat) < TOL
    assert helpers.rel_err(want, model.image_prompt_feat) < TOL
