"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU fp32 restatement (plain torch functional ops, autograd for gradients) of the reference's ALPRO forward path, used
only as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. Nothing in
alpro_b200/ may import this module.

Parity status: PINNED. oracle/make_golden.py imports the unmodified reference from /root/reference in the build
container, runs it on seeded synthetic inputs and writes tests/golden/*.npz; tests/test_oracle_golden.py checks this
restatement against those vectors (and, when /root/reference is present, against the live reference).
The Horovod collective semantics (hvd.allgather fwd = concat in rank order, bwd = sum-allreduce + narrow; gradient
averaging) come from horovod==0.19.4, which is not vendored in the reference: that part is "parity unpinned" by any
reference test and is pinned only by our own world-size-2 emulation (tests/test_distributed_cpu.py).

Every function cites the reference file:line it follows (paths relative to /root/reference).
State dicts use the reference's parameter names, so the same dict loads into the reference model,
this oracle and the CUDA model.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------
# TimeSformer (divided space-time)            src/modeling/timesformer/vit.py
# --------------------------------------------------------------------------------------------------------------
def _mha(x, wqkv, bqkv, wproj, bproj, heads):
    """Attention.forward, vit.py:81-100: qkv Linear -> [3,B,h,N,dh]; softmax(q k^T * dh^-0.5) v; proj."""
    Bp, S, C = x.shape
    dh = C // heads
    qkv = F.linear(x, wqkv, bqkv).view(Bp, S, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = torch.softmax((q @ k.transpose(-1, -2)) * (dh ** -0.5), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(Bp, S, C)
    return F.linear(o, wproj, bproj)


def vit_block(sd, p, x, B, T, heads, eps=1e-6, dp=None):
    """Block.forward (divided_space_time), vit.py:136-213. dp=None: eval mode (drop_path = identity); otherwise a dict
    of per-sample DropPath factors mask/keep_prob (vit_utils.py:137-162) injected by the test: m_t [B,N] (temporal
    branch, mask shape (b h w,1,1)), m_s [B,T] (spatial, (b t,1,1)), m_m [B] (MLP, (b,1,1)).
    x: [B, 1 + N*T, d] with token index 1 + n*T + t ('b (h w t) m', vit.py:147)."""
    d = x.shape[-1]
    N = (x.shape[1] - 1) // T
    g = lambda n: sd[p + n]
    # temporal attention over the T tokens of each patch position (vit.py:146-162)
    xt = x[:, 1:, :].reshape(B * N, T, d)
    t_in = F.layer_norm(xt, (d,), g("temporal_norm1.weight"), g("temporal_norm1.bias"), eps)
    t_out = _mha(t_in, g("temporal_attn.qkv.weight"), g("temporal_attn.qkv.bias"), g("temporal_attn.proj.weight"),
                 g("temporal_attn.proj.bias"), heads)
    if dp is not None:
        t_out = t_out * dp["m_t"].reshape(B * N, 1, 1)                    # res_temporal = drop_path(...), vit.py:157
    t_out = F.linear(t_out.reshape(B, N * T, d), g("temporal_fc.weight"), g("temporal_fc.bias"))
    xt = x[:, 1:, :] + t_out
    # spatial attention over cls + N patches of each frame (vit.py:165-191)
    cls0 = x[:, :1, :]                                             # [B,1,d]
    xs = xt.reshape(B, N, T, d).permute(0, 2, 1, 3).reshape(B * T, N, d)
    cls_rep = cls0.expand(B, T, d).reshape(B * T, 1, d)
    xs = torch.cat([cls_rep, xs], dim=1)
    s_in = F.layer_norm(xs, (d,), g("norm1.weight"), g("norm1.bias"), eps)
    s_out = _mha(s_in, g("attn.qkv.weight"), g("attn.qkv.bias"), g("attn.proj.weight"), g("attn.proj.bias"), heads)
    if dp is not None:
        s_out = s_out * dp["m_s"].reshape(B * T, 1, 1)                    # res_spatial = drop_path(...), vit.py:181
    cls_out = s_out[:, 0, :].reshape(B, T, d).mean(dim=1, keepdim=True)   # vit.py:184-187
    res = s_out[:, 1:, :].reshape(B, T, N, d).permute(0, 2, 1, 3).reshape(B, N * T, d)
    x = torch.cat([cls0, xt], dim=1) + torch.cat([cls_out, res], dim=1)    # vit.py:195-196
    # MLP (vit.py:198-212, Mlp.forward :59-65)
    m_in = F.layer_norm(x, (d,), g("norm2.weight"), g("norm2.bias"), eps)
    hdn = F.gelu(F.linear(m_in, g("mlp.fc1.weight"), g("mlp.fc1.bias")))
    m_out = F.linear(hdn, g("mlp.fc2.weight"), g("mlp.fc2.bias"))
    if dp is not None:
        m_out = m_out * dp["m_m"].reshape(B, 1, 1)                        # x_res + drop_path(mlp_out), vit.py:212
    return x + m_out


def vit_tokens(sd, p, frames, patch):
    """PatchEmbed.forward vit.py:233-239 + embedding part of VisionTransformer.forward_features vit.py:321-361.
    frames: [B,T,3,H,W] (the task models transpose to b c t h w before calling; alpro_models.py:188-190).
    Returns x [B, 1+N*T, d] in the (n t) token order."""
    B, T, C, H, W = frames.shape
    w, b = sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"]
    d = w.shape[0]
    y = F.conv2d(frames.reshape(B * T, C, H, W), w, b, stride=patch)      # [(b t), d, gh, gw]
    gh, gw = y.shape[-2:]
    N = gh * gw
    y = y.flatten(2).transpose(1, 2)                                       # [(b t), N, d]
    pos = sd[p + "pos_embed"]                                              # [1, 1+P, d]
    if pos.shape[1] != N + 1:                                              # nearest resize, vit.py:328-340
        P = int(round(math.sqrt(pos.shape[1] - 1)))
        grid = pos[0, 1:].t().reshape(1, d, P, P)
        grid = F.interpolate(grid, size=(gh, gw), mode="nearest").flatten(2).transpose(1, 2)
        pos = torch.cat([pos[:, :1], grid], dim=1)
    cls = sd[p + "cls_token"] + pos[:, :1]                                 # cls gets pos_embed[0] only (vit.py:342,345)
    y = y + pos[:, 1:]
    tim = sd[p + "time_embed"]                                             # [1, T0, d]
    if tim.shape[1] != T:                                                  # vit.py:350-355
        tim = F.interpolate(tim.transpose(1, 2), size=T, mode="nearest").transpose(1, 2)
    y = y.reshape(B, T, N, d) + tim.reshape(1, T, 1, d)                    # (b n) t m + time_embed, vit.py:348-357
    y = y.permute(0, 2, 1, 3).reshape(B, N * T, d)                         # 'b (n t) m', vit.py:359
    return torch.cat([cls.expand(B, 1, d), y], dim=1)


class RandomDrop:
    """Train-mode regularisers drawn by torch at run time (timing legs of bench.py; parity tests inject masks instead):
    hidden dropout p_hidden (xbert.py:178,358,436), attention-probability dropout p_attn (xbert.py:331)."""

    def __init__(self, p_hidden, p_attn):
        self.p_hidden, self.p_attn = p_hidden, p_attn


def random_train(cfg, vis, drop_path_rate=0.1):
    """`train=` argument that makes retrieval_forward / pretrain_forward draw their own masks (reference .train())."""
    rd = {i: RandomDrop(cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"])
          for i in range(cfg["num_hidden_layers"])}
    return dict(drop_path=float(drop_path_rate), emb=float(cfg["hidden_dropout_prob"]), text=rd, pos=rd, neg=rd,
                emb_mlm=float(cfg["hidden_dropout_prob"]), text_mlm=rd, mlm=rd)


def _random_drop_path(rate, depth, B, N, T, dev):
    """DropPath draws of every block (vit_utils.py:137-162; rate scaled linearly over depth, vit.py:272-277)."""
    out = []
    for i in range(depth):
        pr = rate * i / max(depth - 1, 1)
        if pr <= 0:
            out.append(None)
            continue
        keep = 1.0 - pr
        draw = lambda *shape: torch.bernoulli(torch.full(shape, keep, device=dev)) / keep
        out.append(dict(m_t=draw(B, N), m_s=draw(B, T), m_m=draw(B)))
    return out


def visual_forward(sd, p, frames, vis, return_tokens=False, drop_path=None):
    """TimeSformer.forward_features vit.py:475-503 (pooling='temporal'): blocks, final LN (vit.py:372), mean over t.
    p is the prefix of the VisionTransformer ('visual_encoder.model.'). Returns [B, 1+N, d]."""
    B, T = frames.shape[:2]
    x = vit_tokens(sd, p, frames, vis["patch"])
    if isinstance(drop_path, float):
        drop_path = _random_drop_path(drop_path, vis["depth"], B, (x.shape[1] - 1) // T, T, x.device)
    for i in range(vis["depth"]):
        x = vit_block(sd, f"{p}blocks.{i}.", x, B, T, vis["heads"], dp=drop_path[i] if drop_path else None)
    d = x.shape[-1]
    x = F.layer_norm(x, (d,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    N = (x.shape[1] - 1) // T
    pooled = x[:, 1:].reshape(B, N, T, d).mean(dim=2)
    out = torch.cat([x[:, :1], pooled], dim=1)
    return (out, x) if return_tokens else out


# --------------------------------------------------------------------------------------------------------------
# BERT (text / fusion modes)                    src/modeling/xbert.py
# --------------------------------------------------------------------------------------------------------------
def bert_embeddings(sd, p, ids, eps, mask=None):
    """BertEmbeddings.forward xbert.py:186-213 (token_type all zero, absolute positions). mask: injected dropout
    mask/keep for train-mode parity (None = eval)."""
    e = p + "bert.embeddings."
    L = ids.shape[1]
    x = sd[e + "word_embeddings.weight"][ids] + sd[e + "token_type_embeddings.weight"][0] \
        + sd[e + "position_embeddings.weight"][:L].unsqueeze(0)
    h = x.shape[-1]
    y = F.layer_norm(x, (h,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    if isinstance(mask, float):
        return F.dropout(y, mask, True)
    return y if mask is None else y * mask.reshape(y.shape)


def bert_layer(sd, l, x, ext_mask, heads, eps, drop=None):
    """BertLayer.forward xbert.py:457-519 = BertSelfAttention :263-346 + BertSelfOutput :349-360 +
    BertIntermediate :412-424 + BertOutput :427-438 (post-LN; has_cross_attention=False :450)."""
    Bp, S, h = x.shape
    dh = h // heads
    def proj(n):
        return F.linear(x, sd[l + f"attention.self.{n}.weight"], sd[l + f"attention.self.{n}.bias"]) \
            .view(Bp, S, heads, dh).transpose(1, 2)
    q, k, v = proj("query"), proj("key"), proj("value")
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(dh) + ext_mask           # xbert.py:317-320
    probs = torch.softmax(scores, dim=-1)
    rnd = isinstance(drop, RandomDrop)
    if rnd:
        probs = F.dropout(probs, drop.p_attn, True)
    elif drop is not None and len(drop) > 2 and drop[2] is not None:
        probs = probs * drop[2]                                             # attention_probs dropout, xbert.py:331
    ctx = (probs @ v).transpose(1, 2).reshape(Bp, S, h)
    a = F.linear(ctx, sd[l + "attention.output.dense.weight"], sd[l + "attention.output.dense.bias"])
    if rnd:
        a = F.dropout(a, drop.p_hidden, True)
    elif drop is not None:
        a = a * drop[0].reshape(a.shape)                                    # BertSelfOutput.dropout, xbert.py:358
    a = F.layer_norm(a + x, (h,), sd[l + "attention.output.LayerNorm.weight"], sd[l + "attention.output.LayerNorm.bias"], eps)
    i = F.gelu(F.linear(a, sd[l + "intermediate.dense.weight"], sd[l + "intermediate.dense.bias"]))
    o = F.linear(i, sd[l + "output.dense.weight"], sd[l + "output.dense.bias"])
    if rnd:
        o = F.dropout(o, drop.p_hidden, True)
    elif drop is not None:
        o = o * drop[1].reshape(o.shape)                                    # BertOutput.dropout, xbert.py:436
    return F.layer_norm(o + a, (h,), sd[l + "output.LayerNorm.weight"], sd[l + "output.LayerNorm.bias"], eps)


def bert_encode(sd, p, x, mask, cfg, mode, drops=None):
    """BertModel.forward xbert.py:940-1081 + BertEncoder.forward :528-630: extended mask (1-mask)*-10000
    (get_extended_attention_mask :878-938), layers [0,fusion_layer) for 'text', [fusion_layer,L) for 'fusion'."""
    ext = (1.0 - mask.to(torch.float32))[:, None, None, :] * -10000.0
    lo, hi = (0, cfg["fusion_layer"]) if mode == "text" else (cfg["fusion_layer"], cfg["num_hidden_layers"])
    for i in range(lo, hi):
        x = bert_layer(sd, f"{p}bert.encoder.layer.{i}.", x, ext, cfg["num_attention_heads"], cfg["layer_norm_eps"],
                       drop=drops[i] if drops else None)
    return x


def bert_text(sd, p, ids, mask, cfg, emb_mask=None, drops=None):
    return bert_encode(sd, p, bert_embeddings(sd, p, ids, cfg["layer_norm_eps"], emb_mask), mask, cfg, "text", drops)


def mlm_head(sd, p, x, eps):
    """BertOnlyMLMHead -> BertLMPredictionHead xbert.py:648-692: dense, gelu, LN, tied decoder + bias."""
    c = p + "cls.predictions."
    h = x.shape[-1]
    t = F.gelu(F.linear(x, sd[c + "transform.dense.weight"], sd[c + "transform.dense.bias"]))
    t = F.layer_norm(t, (h,), sd[c + "transform.LayerNorm.weight"], sd[c + "transform.LayerNorm.bias"], eps)
    return F.linear(t, sd[p + "bert.embeddings.word_embeddings.weight"], sd[c + "bias"])


# --------------------------------------------------------------------------------------------------------------
# Task heads                                     src/modeling/alpro_models.py
# --------------------------------------------------------------------------------------------------------------
def argmax_sampler(weights_row):
    """Deterministic stand-in for torch.multinomial(w, 1).item() (alpro_models.py:303,310,835,842) used on both
    sides of every parity comparison."""
    return int(torch.argmax(weights_row).item())


def vtc(sd, pfx, video_cls, text_cls, rank=0, gather=None):
    """VTC / 'itc' (alpro_models.py:103-128, 750-779): proj + L2 normalise, all-gather, sims / temp, two CEs."""
    temp = sd[pfx + "temp"].clamp(0.001, 0.5)                              # temp.clamp_ :80-81, :734-735
    vf = F.normalize(F.linear(video_cls, sd[pfx + "vision_proj.weight"], sd[pfx + "vision_proj.bias"]), dim=-1)
    tf = F.normalize(F.linear(text_cls, sd[pfx + "text_proj.weight"], sd[pfx + "text_proj.bias"]), dim=-1)
    gv = gather(vf) if gather else vf
    gt = gather(tf) if gather else tf
    sim_v2t = vf @ gt.t() / temp
    sim_t2v = tf @ gv.t() / temp
    b = vf.shape[0]
    tgt = torch.arange(b, device=vf.device) + b * rank                     # sim_targets block, :119-123
    loss = 0.5 * (F.cross_entropy(sim_v2t, tgt) + F.cross_entropy(sim_t2v, tgt))
    return loss, sim_v2t, sim_t2v, vf, tf


def mine_negatives(sim_v2t, sim_t2v, rank, sampler):
    """Hard-negative indices (alpro_models.py:288-316, 819-847): softmax of the local [b,b] block with -inf diagonal;
    first b draws pick a negative *video* per text (weights_t2v rows), next b draws a negative *text* per video."""
    b = sim_v2t.shape[0]
    with torch.no_grad():
        blk = slice(b * rank, b * (rank + 1))
        w_v2t = sim_v2t[:, blk].clone().fill_diagonal_(-float("inf")).softmax(dim=1)
        w_t2v = sim_t2v[:, blk].clone().fill_diagonal_(-float("inf")).softmax(dim=1)
    neg_video = [sampler(w_t2v[i]) for i in range(b)]
    neg_text = [sampler(w_v2t[i]) for i in range(b)]
    return neg_video, neg_text


def vtm(sd, pfx, cfg, text_embeds, text_mask, video_embeds, neg_video, neg_text, drops_pos=None, drops_neg=None):
    """compute_vtm (alpro_models.py:269-344, 800-872): fusion encoder on positives (text_i, video_i), then on
    (text_i, video_neg_i) and (text_neg_i, video_i); itm_head on the [CLS] outputs; CE with labels [1]*b + [0]*2b.
    Returns loss, logits [3b,2], labels, positive fusion output [b, L+1+N, h]."""
    b, L = text_mask.shape
    nv = video_embeds.shape[1]
    dev = video_embeds.device
    ones = torch.ones(b, nv, dtype=text_mask.dtype, device=dev)
    nvi = torch.tensor(neg_video, dtype=torch.long, device=dev)
    nti = torch.tensor(neg_text, dtype=torch.long, device=dev)
    emb_pos = torch.cat([text_embeds, video_embeds], dim=1)
    mask_pos = torch.cat([text_mask, ones], dim=1)
    out_pos = bert_encode(sd, pfx + "text_encoder.", emb_pos, mask_pos, cfg, "fusion", drops_pos)
    txt_all = torch.cat([text_embeds, text_embeds[nti]], dim=0)
    msk_all = torch.cat([text_mask, text_mask[nti]], dim=0)
    vid_all = torch.cat([video_embeds[nvi], video_embeds], dim=0)
    emb_neg = torch.cat([txt_all, vid_all], dim=1)
    mask_neg = torch.cat([msk_all, torch.cat([ones, ones], dim=0)], dim=1)
    out_neg = bert_encode(sd, pfx + "text_encoder.", emb_neg, mask_neg, cfg, "fusion", drops_neg)
    cls = torch.cat([out_pos[:, 0], out_neg[:, 0]], dim=0)
    logits = F.linear(cls, sd[pfx + "itm_head.weight"], sd[pfx + "itm_head.bias"])
    labels = torch.cat([torch.ones(b, dtype=torch.long, device=dev), torch.zeros(2 * b, dtype=torch.long, device=dev)])
    return F.cross_entropy(logits, labels), logits, labels, out_pos


def retrieval_forward(sd, cfg, vis, batch, rank=0, gather=None, sampler=argmax_sampler, train=None):
    """AlproForVideoTextRetrieval.forward alpro_models.py:733-798. `train` (tests only): injected regulariser masks
    dict(drop_path=[per-block dict], emb=mask, text={layer: (mo, mf)}, pos={layer: ...}, neg={layer: ...})."""
    tr = train or {}
    video_embeds = visual_forward(sd, "visual_encoder.model.", batch["visual_inputs"], vis, drop_path=tr.get("drop_path"))
    text_embeds = bert_text(sd, "text_encoder.", batch["text_input_ids"], batch["text_input_mask"], cfg,
                            tr.get("emb"), tr.get("text"))
    itc_loss, s_v2t, s_t2v, vf, tf = vtc(sd, "", video_embeds[:, 0], text_embeds[:, 0], rank, gather)
    neg_v, neg_t = mine_negatives(s_v2t.detach(), s_t2v.detach(), rank, sampler)
    itm_loss, itm_scores, itm_labels, _ = vtm(sd, "", cfg, text_embeds, batch["text_input_mask"], video_embeds,
                                              neg_v, neg_t, tr.get("pos"), tr.get("neg"))
    return dict(itm_scores=itm_scores, itm_loss=itm_loss, itm_labels=itm_labels, itc_loss=itc_loss,
                _video_embeds=video_embeds, _text_embeds=text_embeds, _neg_video=neg_v, _neg_text=neg_t,
                _video_feat=vf, _text_feat=tf)


def inference_forward(sd, cfg, vis, batch):
    """AlproForVideoTextRetrieval.forward_inference alpro_models.py:874-914 (1 video x n texts)."""
    video_embeds = visual_forward(sd, "visual_encoder.model.", batch["visual_inputs"], vis)
    text_embeds = bert_text(sd, "text_encoder.", batch["text_input_ids"], batch["text_input_mask"], cfg)
    temp = sd["temp"]
    vf = F.normalize(F.linear(video_embeds[:, 0], sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1)
    tf = F.normalize(F.linear(text_embeds[:, 0], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1)
    n = text_embeds.shape[0]
    ve = video_embeds.repeat(n, 1, 1)
    mask = torch.cat([batch["text_input_mask"], torch.ones(ve.shape[:2], dtype=torch.long, device=ve.device)], dim=1)
    out = bert_encode(sd, "text_encoder.", torch.cat([text_embeds, ve], dim=1), mask, cfg, "fusion")
    logits = F.linear(out[:, 0], sd["itm_head.weight"], sd["itm_head.bias"])
    return dict(logits=logits, itc_scores=vf @ tf.t() / temp)


def pseudo_labels(sd, cfg, vis, batch):
    """Prompter.get_pseudo_labels alpro_models.py:531-551 + _compute_soft_labels :525-529 (teacher, no grad).
    NB bug-compatible: ignore iff argmax *index* < 0.2, i.e. iff the top entity is index 0."""
    with torch.no_grad():
        ve = visual_forward(sd, "prompter.visual_encoder.model.", batch["crop_visual_inputs"], vis)
        feat = F.normalize(F.linear(ve[:, 0], sd["prompter.vision_proj.weight"], sd["prompter.vision_proj.bias"]), dim=-1)
        prompt = sd["prompter.video_prompt_feat"] if batch.get("type", "video") == "video" else sd["prompter.image_prompt_feat"]
        sim = feat @ prompt.t() / sd["prompter.temp"]
        soft = torch.softmax(sim, dim=1)
        ignore = torch.max(sim, dim=1)[1] < 0.2
    return soft, ignore


def pretrain_forward(sd, cfg, vis, batch, rank=0, gather=None, sampler=argmax_sampler, train=None):
    """AlproForPretrain.forward alpro_models.py:79-183 (use_mask_prob = 0 so context_visual_inputs is unused).
    `train` (tests only): injected regulariser masks as in retrieval_forward, plus the MLM branch's own draws:
    emb_mlm / text_mlm (text pass over the masked ids, compute_mlm :352-356) and mlm (its fusion pass :358-364).
    The teacher (Prompter.get_pseudo_labels :531-535) always runs in eval mode."""
    pfx = ""
    tr = train or {}
    video_embeds = visual_forward(sd, "visual_encoder.model.", batch["visual_inputs"], vis, drop_path=tr.get("drop_path"))
    mask = batch["text_input_mask"]
    text_embeds = bert_text(sd, "text_encoder.", batch["text_input_ids"], mask, cfg, tr.get("emb"), tr.get("text"))
    itc_loss, s_v2t, s_t2v, vf, tf = vtc(sd, pfx, video_embeds[:, 0], text_embeds[:, 0], rank, gather)
    neg_v, neg_t = mine_negatives(s_v2t.detach(), s_t2v.detach(), rank, sampler)
    itm_loss, itm_scores, itm_labels, out_pos = vtm(sd, pfx, cfg, text_embeds, mask, video_embeds, neg_v, neg_t,
                                                    tr.get("pos"), tr.get("neg"))
    out = dict(itc_loss=itc_loss, itm_scores=itm_scores, itm_loss=itm_loss, itm_labels=itm_labels,
               mlm_scores=None, mlm_loss=None, mlm_labels=None, mpm_loss=None, mpm_logits=None, mpm_labels=None,
               _video_embeds=video_embeds, _text_embeds=text_embeds, _neg_video=neg_v, _neg_text=neg_t)
    b, L = mask.shape
    nv = video_embeds.shape[1]
    ones = torch.ones(b, nv, dtype=mask.dtype, device=mask.device)
    if "mlm_labels" in batch:                                               # compute_mlm :346-373
        mt = bert_text(sd, "text_encoder.", batch["mlm_text_input_ids"], mask, cfg, tr.get("emb_mlm"), tr.get("text_mlm"))
        fo = bert_encode(sd, "text_encoder.", torch.cat([mt, video_embeds], dim=1), torch.cat([mask, ones], dim=1),
                         cfg, "fusion", tr.get("mlm"))
        scores = mlm_head(sd, "text_encoder.", fo[:, :L], cfg["layer_norm_eps"])
        out["mlm_scores"] = scores
        out["mlm_labels"] = batch["mlm_labels"]
        out["mlm_loss"] = F.cross_entropy(scores.reshape(-1, scores.shape[-1]), batch["mlm_labels"].reshape(-1),
                                          ignore_index=-100)
    if "mpm_mask" in batch:                                                 # compute_mpm_with_encoder_out :209-232
        soft, ignore = pseudo_labels(sd, cfg, vis, batch)
        vis_out = out_pos[:, L + 1:]
        inv = (1.0 - batch["mpm_mask"].reshape(b, -1)).unsqueeze(-1)
        pooled = (inv * vis_out).sum(dim=1) / inv.squeeze(-1).sum(dim=-1, keepdim=True)
        hdn = F.relu(F.linear(pooled, sd["mpm_head.0.weight"], sd["mpm_head.0.bias"]))
        logits = F.linear(hdn, sd["mpm_head.2.weight"], sd["mpm_head.2.bias"])
        ce = -(F.log_softmax(logits, dim=1) * soft).sum(dim=1)
        ce = torch.where(ignore, torch.zeros_like(ce), ce)
        out["mpm_loss"] = ce.sum() / (b - ignore.sum())
        out["mpm_logits"] = logits
        out["mpm_labels"] = soft
        out["_mpm_ignore"] = ignore
    return out


# --------------------------------------------------------------------------------------------------------------
# Prompter (teacher)                             src/modeling/alpro_models.py:389-630
# --------------------------------------------------------------------------------------------------------------
def prompter_forward(sd, cfg, vis, batch, rank=0, gather=None):
    """Prompter.forward alpro_models.py:553-595 (+ forward_feats :597-630): the contrastive objective alone, with the
    reference's soft-target formulation -sum(log_softmax(sim) * targets).mean(); returns its four outputs.
    `sd` holds the Prompter's own state dict (no 'prompter.' prefix)."""
    temp = sd["temp"].clamp(0.001, 0.5)
    video_embeds = visual_forward(sd, "visual_encoder.model.", batch["visual_inputs"], vis)
    text_embeds = bert_text(sd, "text_encoder.", batch["text_input_ids"], batch["text_input_mask"], cfg)
    vf = F.normalize(F.linear(video_embeds[:, 0], sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1)
    tf = F.normalize(F.linear(text_embeds[:, 0], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1)
    gv = gather(vf) if gather else vf
    gt = gather(tf) if gather else tf
    sim_v2t = vf @ gt.t() / temp
    sim_t2v = tf @ gv.t() / temp
    b = vf.shape[0]
    targets = torch.zeros_like(sim_v2t)
    targets[:, b * rank:b * (rank + 1)] = torch.eye(b, device=vf.device)
    lv, lt = F.log_softmax(sim_v2t, dim=1), F.log_softmax(sim_t2v, dim=1)
    loss = 0.5 * (-(lv * targets).sum(dim=1).mean() - (lt * targets).sum(dim=1).mean())
    return dict(itc_loss=loss, itc_labels=targets.max(dim=1)[1], i2t_scores=lv, t2i_scores=lt,
                _video_embeds=video_embeds, _video_feat=vf, _text_embeds=text_embeds, _text_feat=tf)


def build_text_prompts(sd, cfg, ids, mask, num_entities):
    """Prompter.build_text_prompts alpro_models.py:430-507 for one prompt set: text-encode every prompt, project and
    normalise the [CLS] output, then average over templates — prompts are laid out template-major
    (chunk(num_templates) then stack(dim=1), :470-474)."""
    with torch.no_grad():
        te = bert_text(sd, "text_encoder.", ids, mask, cfg)
        feat = F.normalize(F.linear(te[:, 0], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1)
        n_templates = int(feat.shape[0] / num_entities)
        return torch.stack(feat.chunk(n_templates), dim=1).mean(dim=1)
