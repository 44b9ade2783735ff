"""2-GPU test (NCCL): data-parallel retrieval step through alpro_b200.comm against a single-process oracle statement of
the same global objective (mean over ranks of the per-rank losses; gradient = average). Skipped with < 2 GPUs.
Run on the box:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from alpro_b200 import comm
    from oracle import configs
    from tests import helpers
    from tests.test_gpu_parity import build_cuda_model, to_cuda
    cfg = dict(configs.GOLDEN["tiny_retrieval"])
    spec, sd, _ = helpers.make_inputs(cfg)
    from alpro_b200 import synth
    batch = synth.synth_batch("retrieval", cfg["B"], cfg["T"], cfg["img"], cfg["L"], cfg["bert"]["vocab_size"],
                              seed=100 + rank)
    model = build_cuda_model(cfg, sd)
    comm.attach(model)
    out = model(to_cuda(batch))
    (out["itc_loss"] + out["itm_loss"]).backward()
    comm.allreduce_gradients(model)
    torch.cuda.synchronize()
    res = {"itc": float(out["itc_loss"]), "itm": float(out["itm_loss"]),
           "neg": out["_neg_video"].tolist() + out["_neg_text"].tolist()}
    if rank == 0:
        res["grads"] = {n: p.grad.float().cpu() for n, p in model.named_parameters() if p.grad is not None}
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_global_oracle():
    import torch.multiprocessing as mp
    from alpro_b200 import synth
    from oracle import alpro_oracle, configs
    from tests import helpers
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process oracle of the global objective
    cfg = configs.GOLDEN["tiny_retrieval"]
    spec, sd, _ = helpers.make_inputs(cfg)
    sd = {k: v.clone() for k, v in sd.items()}
    for k in list(sd):
        c = synth.canonical_name(k)
        if c != k:
            sd[k] = sd[c]
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    batches = [synth.synth_batch("retrieval", cfg["B"], cfg["T"], cfg["img"], cfg["L"], cfg["bert"]["vocab_size"],
                                 seed=100 + r) for r in range(world)]
    ve = [alpro_oracle.visual_forward(sd, "visual_encoder.model.", b["visual_inputs"], cfg["vis"]) for b in batches]
    te = [alpro_oracle.bert_text(sd, "text_encoder.", b["text_input_ids"], b["text_input_mask"], cfg["bert"]) for b in batches]
    import torch.nn.functional as F
    vf = [F.normalize(F.linear(v[:, 0], sd["vision_proj.weight"], sd["vision_proj.bias"]), dim=-1) for v in ve]
    tf = [F.normalize(F.linear(t[:, 0], sd["text_proj.weight"], sd["text_proj.bias"]), dim=-1) for t in te]
    total = 0
    for r in range(world):
        # oracle.vtc gathers the video features first, then the text features: hand it the global (differentiable)
        # feature matrices in that order
        answers = iter([torch.cat(vf), torch.cat(tf)])
        loss, s_v2t, s_t2v, _, _ = alpro_oracle.vtc(sd, "", ve[r][:, 0], te[r][:, 0], r, gather=lambda t: next(answers))
        neg_v, neg_t = alpro_oracle.mine_negatives(s_v2t.detach(), s_t2v.detach(), r, alpro_oracle.argmax_sampler)
        itm, _, _, _ = alpro_oracle.vtm(sd, "", cfg["bert"], te[r], batches[r]["text_input_mask"], ve[r], neg_v, neg_t)
        assert abs(float(loss) - got[r]["itc"]) < 1e-3 * max(1.0, abs(float(loss)))
        assert abs(float(itm) - got[r]["itm"]) < 1e-3 * max(1.0, abs(float(itm)))
        assert neg_v + neg_t == got[r]["neg"]
        total = total + (loss + itm) / world
    total.backward()
    bad = []
    gmax = max(float(v.grad.abs().max()) for v in sd.values() if v.grad is not None)
    for n, g in got[0]["grads"].items():
        ref = sd[n].grad
        if ref is None or float(ref.abs().max()) < 1e-6 * gmax:
            continue
        e = helpers.rel_err(g, ref)
        if e > 3e-2:
            bad.append((n, e))
    assert not bad, bad[:8]
