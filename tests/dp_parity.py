"""Data-parallel parity check shared by tests/test_gpu_multi.py and bench.py (WORLD_SIZE > 1, outside the timed region):
every rank runs one retrieval training step of the tiny parity config on ITS OWN batch through alpro_b200.comm (VTC
feature exchange + gradient averaging); rank 0 restates the same global objective with the CPU oracle (mean over ranks
of the per-rank losses, features gathered in rank order) and compares losses, hard-negative indices and every averaged
parameter gradient. TEST INFRASTRUCTURE: the oracle is the checker only."""
import torch
import torch.distributed as dist


def run(device, world, rank):
    """Collective: call on every rank of an initialised process group. Returns the verdict dict on rank 0, None elsewhere."""
    import torch.nn.functional as F
    from alpro_b200 import comm, synth
    from oracle import alpro_oracle, configs
    from tests import helpers
    from tests.test_gpu_parity import build_cuda_model, to_cuda
    cfg = dict(configs.GOLDEN["tiny_retrieval"])
    spec, sd, _ = helpers.make_inputs(cfg)
    mk = lambda r: synth.synth_batch("retrieval", cfg["B"], cfg["T"], cfg["img"], cfg["L"], cfg["bert"]["vocab_size"],
                                     seed=100 + r)
    model = build_cuda_model(cfg, sd)
    comm.attach(model)
    out = model(to_cuda(mk(rank)))
    (out["itc_loss"] + out["itm_loss"]).backward()
    comm.allreduce_gradients(model)
    torch.cuda.synchronize()
    mine = dict(itc=float(out["itc_loss"].detach()), itm=float(out["itm_loss"].detach()),
                neg=out["_neg_video"].tolist() + out["_neg_text"].tolist())
    got = [None] * world
    dist.all_gather_object(got, mine)
    # averaged gradients must be IDENTICAL on every rank
    flat = model.engine.last_grads.flat
    ref0 = flat.clone()
    dist.broadcast(ref0, src=0)
    diff = (flat - ref0).abs().max().reshape(1)
    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    grads = {n: p.grad.float().cpu() for n, p in model.named_parameters() if p.grad is not None}
    # single-process oracle of the global objective
    sd = {k: v.clone() for k, v in sd.items()}
    for k in list(sd):
        c = synth.canonical_name(k)
        if c != k:
            sd[k] = sd[c]
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    batches = [mk(r) for r in range(world)]
    ve = [alpro_oracle.visual_forward(sd, "visual_encoder.model.", b["visual_inputs"], cfg["vis"]) for b in batches]
    te = [alpro_oracle.bert_text(sd, "text_encoder.", b["text_input_ids"], b["text_input_mask"], cfg["bert"]) for b in batches]
    vf = [F.normalize(F.linear(v[:, 0], sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1) for v in ve]
    tf = [F.normalize(F.linear(t[:, 0], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1) for t in te]
    total, max_rel_loss, neg_ok = 0, 0.0, True
    for r in range(world):
        # oracle.vtc gathers the video features first, then the text features: hand it the global (differentiable)
        # feature matrices in that order
        answers = iter([torch.cat(vf), torch.cat(tf)])
        loss, s_v2t, s_t2v, _, _ = alpro_oracle.vtc(sd, "", ve[r][:, 0], te[r][:, 0], r, gather=lambda t: next(answers))
        neg_v, neg_t = alpro_oracle.mine_negatives(s_v2t.detach(), s_t2v.detach(), r, alpro_oracle.argmax_sampler)
        itm, _, _, _ = alpro_oracle.vtm(sd, "", cfg["bert"], te[r], batches[r]["text_input_mask"], ve[r], neg_v, neg_t)
        max_rel_loss = max(max_rel_loss, abs(float(loss) - got[r]["itc"]) / max(1e-6, abs(float(loss))),
                           abs(float(itm) - got[r]["itm"]) / max(1e-6, abs(float(itm))))
        neg_ok = neg_ok and (neg_v + neg_t == got[r]["neg"])
        total = total + (loss + itm) / world
    total.backward()
    gmax = max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    max_rel_grad, worst = 0.0, ""
    for n, g in grads.items():
        ref = sd[n].grad
        if ref is None or float(ref.abs().max()) < 1e-6 * gmax:
            continue
        e = helpers.rel_err(g, ref)
        if e > max_rel_grad:
            max_rel_grad, worst = e, n
    return dict(world=world, max_rel_loss=float(f"{max_rel_loss:.3e}"), max_rel_grad=float(f"{max_rel_grad:.3e}"),
                worst_grad=worst, neg_indices_equal=bool(neg_ok), grads_identical_across_ranks=float(diff) == 0.0,
                config="tiny_retrieval (d=192, 2 blocks, 2+2 BERT layers), one batch of 3 pairs per rank, vs CPU oracle "
                       "of the global objective")
