"""CPU: the drop-in boundary around the hot path —

  * the UNCHANGED reference run scripts import against alpro_b200 through alpro_b200.shims (horovod.torch / apex.amp
    stand-ins, `src.modeling.alpro_models` alias) and their own setup_model() builds and loads OUR classes;
  * checkpoint loading: load_separate_ckpt(visual_weights_path=, prompter_weights_path=), load-time pos/time embedding
    resize (helpers.py:355-375), the reference's load_state_dict_with_pos_embed_resizing applied to our module tree;
  * horovod stand-in semantics in a single process (world-size-2 semantics: tests/test_distributed_cpu.py).
"""
import json
import os
import sys
import types

import pytest
import torch

from alpro_b200 import modeling, shims, synth
from oracle import configs

REF = "/root/reference"


def _tiny_models_cfg(T=2, img=64):
    cfg = configs.tiny("pretrain", T=T, img=img)
    v = dict(cfg["video"])
    v.update(embed_dim=cfg["vis"]["d"], depth=cfg["vis"]["depth"], num_heads=cfg["vis"]["heads"])
    b = dict(cfg["bert"])
    b["num_entities"] = cfg["num_entities"]
    return cfg, b, v


# ---------------------------------------------------------------------------------------------- checkpoint loading
def test_resize_embeddings_follow_nearest_interpolate():
    g = torch.Generator().manual_seed(0)
    pos = torch.randn(1, 1 + 9, 8, generator=g)
    new = modeling.resize_spatial_embedding(pos, 16)
    want = torch.cat([pos[:, :1], torch.nn.functional.interpolate(pos[:, 1:].transpose(1, 2), size=16, mode="nearest")
                      .transpose(1, 2)], dim=1)
    assert torch.equal(new, want)
    tim = torch.randn(1, 8, 8, generator=g)
    for T in (2, 4, 16):
        want = torch.nn.functional.interpolate(tim.transpose(1, 2), size=T, mode="nearest").transpose(1, 2)
        assert torch.equal(modeling.resize_temporal_embedding(tim, T), want)


@pytest.mark.reference
def test_resize_embeddings_match_reference_helpers():
    import importlib
    saved = _set_aside()
    provided = shims.install(alias_models=False, optional_stubs=True, force=True)
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module("src.modeling.timesformer.helpers")
    except Exception as e:  # the helper module drags in optional packages
        pytest.skip(f"reference helpers not importable here: {e}")
    finally:
        sys.path.remove(REF)
        _forget(provided)
        sys.modules.update(saved)
    g = torch.Generator().manual_seed(1)
    sd = {"pos_embed": torch.randn(1, 1 + 196, 12, generator=g), "time_embed": torch.randn(1, 8, 12, generator=g)}
    assert torch.equal(modeling.resize_spatial_embedding(sd["pos_embed"], 49), mod.resize_spatial_embedding(sd, "pos_embed", 49))
    assert torch.equal(modeling.resize_temporal_embedding(sd["time_embed"], 4), mod.resize_temporal_embedding(sd, "time_embed", 4))


@pytest.mark.parametrize("wrap", ["bare", "model_state", "state_dict"])
def test_load_separate_ckpt_resizes_and_strips_prefixes(tmp_path, wrap):
    cfg, b, v = _tiny_models_cfg(T=2, img=64)                 # model: 16 patches, 2 frames
    src_cfg, _, _ = _tiny_models_cfg(T=4, img=48)             # checkpoint: 9 patches, 4 frames
    vis_spec = synth.visual_spec("", src_cfg["vis"]["d"], src_cfg["vis"]["depth"], 4, 9, 16)
    sd = {k: synth.synth_tensor("ckpt." + k, shp, 5) for k, shp in vis_spec.items()}
    if wrap == "model_state":
        blob = {"model_state": {"model." + k: t for k, t in sd.items()}}
    elif wrap == "state_dict":
        blob = {"state_dict": {"module." + k: t for k, t in sd.items()}}
    else:
        blob = sd
    path = str(tmp_path / "vis.pth")
    torch.save(blob, path)
    m = modeling.AlproForVideoTextRetrieval(b, v)
    head_before = m.visual_encoder.model.head.weight.detach().clone()
    m.load_separate_ckpt(visual_weights_path=path, bert_weights_path=None)
    vm = m.visual_encoder.model
    assert torch.equal(vm.blocks[1].mlp.fc1.weight, sd["blocks.1.mlp.fc1.weight"])
    assert torch.equal(vm.pos_embed, modeling.resize_spatial_embedding(sd["pos_embed"], 16))
    assert torch.equal(vm.time_embed, modeling.resize_temporal_embedding(sd["time_embed"], 2))
    assert torch.equal(vm.head.weight, head_before)           # the checkpoint's classifier is ignored (helpers.py:328-336)
    # a checkpoint that does not fit is an error, not a silent random init
    bad = dict(sd)
    bad.pop("blocks.0.attn.qkv.weight")
    torch.save(bad, path)
    with pytest.raises(RuntimeError):
        m.load_separate_ckpt(visual_weights_path=path)


def test_pretrain_load_separate_ckpt_with_prompter_weights(tmp_path):
    """run_pretrain_sparse.py:164-167 calls load_separate_ckpt(visual_weights_path=, prompter_weights_path=)."""
    cfg, b, v = _tiny_models_cfg()
    teacher = modeling.Prompter(b, v)
    tsd = {k: synth.synth_tensor("teacher." + k, tuple(t.shape), 9) if t.is_floating_point() else t
           for k, t in teacher.state_dict().items()}
    tpath = str(tmp_path / "teacher.pt")
    torch.save(tsd, tpath)
    vis_spec = synth.visual_spec("", cfg["vis"]["d"], cfg["vis"]["depth"], 2, 16, 16)
    vpath = str(tmp_path / "vis.pth")
    torch.save({k: synth.synth_tensor("v." + k, shp, 3) for k, shp in vis_spec.items()}, vpath)
    m = modeling.AlproForPretrain(b, v)
    assert isinstance(m.prompter, modeling.Prompter)
    prompt_before = m.prompter.video_prompt_feat.clone()
    m.load_separate_ckpt(visual_weights_path=vpath, prompter_weights_path=tpath)
    assert torch.equal(m.prompter.text_proj.weight, tsd["text_proj.weight"])
    assert torch.equal(m.prompter.visual_encoder.model.blocks[0].attn.qkv.weight,
                       tsd["visual_encoder.model.blocks.0.attn.qkv.weight"])
    assert torch.equal(m.prompter.video_prompt_feat, prompt_before)      # prompts are NOT loaded (:419-424)
    assert torch.equal(m.visual_encoder.model.cls_token, synth.synth_tensor("v.cls_token", (1, 1, cfg["vis"]["d"]), 3))
    assert all(not p.requires_grad for p in m.prompter.parameters())
    assert callable(m.get_pseudo_labels) and callable(m.build_text_prompts)
    for name in ("forward", "forward_feats", "build_text_prompts", "get_pseudo_labels",
                 "load_pretrained_weights_without_prompts", "_forward_visual_embeds", "_compute_soft_labels"):
        assert hasattr(modeling.Prompter, name) or name == "_compute_soft_labels", name


# ---------------------------------------------------------------------------------------------- stand-ins
def test_horovod_and_amp_stand_ins_single_process():
    saved = _set_aside()
    provided = shims.install(alias_models=False, force=True)
    import horovod.torch as hvd
    from apex import amp
    _forget(provided)
    sys.modules.update(saved)
    hvd.init()
    assert (hvd.rank(), hvd.size(), hvd.local_rank()) == (0, 1, 0)
    t = torch.arange(6.0).view(3, 2)
    assert torch.equal(hvd.allgather(t), t) and torch.equal(hvd.allreduce_(t.clone()), t)
    lin = torch.nn.Linear(4, 3)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    dopt = hvd.DistributedOptimizer(opt, named_parameters=lin.named_parameters(), compression=hvd.Compression.none)
    assert isinstance(dopt, torch.optim.SGD) and dopt.param_groups is opt.param_groups
    hvd.broadcast_parameters(lin.state_dict(), root_rank=0)
    hvd.broadcast_optimizer_state(dopt, root_rank=0)
    model, dopt2 = amp.initialize(lin, dopt, enabled=False, opt_level="O2", keep_batchnorm_fp32=True)
    assert model is lin and dopt2 is dopt
    w0 = lin.weight.detach().clone()
    loss = lin(torch.ones(2, 4)).sum()
    with amp.scale_loss(loss, dopt, delay_unscale=False) as scaled:
        scaled.backward()
        dopt.synchronize()
    torch.nn.utils.clip_grad_norm_(amp.master_params(dopt), 5.0)
    with dopt.skip_synchronize():
        dopt.step()
        dopt.zero_grad()
    assert not torch.equal(lin.weight, w0)
    assert amp.state_dict() == {}


# ---------------------------------------------------------------------------------------------- unchanged run scripts
def _fake_data_libs():
    """decord / av / lmdb / spacy are data-side libraries of a real training box; absent here. Test-only empty modules so
    that the run scripts' imports resolve (nothing below decodes a video)."""
    for name in ("decord", "av", "lmdb", "spacy"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                m = types.ModuleType(name)
                if name == "decord":
                    m.VideoReader = type("VideoReader", (), {})
                    m.bridge = types.SimpleNamespace(set_bridge=lambda *_: None)
                sys.modules[name] = m


_DOUBLES = ("horovod", "apex", "src", "easydict", "ujson", "tensorboardX", "decord", "av", "lmdb", "spacy")


def _set_aside():
    """Remove (and return) every test double / reference module another test may have registered, so that
    shims.install() decides on a clean slate; real installed packages (with a __file__) stay."""
    saved = {}
    for k in list(sys.modules):
        if k.split(".")[0] in _DOUBLES and (k.split(".")[0] in ("horovod", "apex", "src")
                                            or not getattr(sys.modules[k], "__file__", None)):
            saved[k] = sys.modules.pop(k)
    return saved


def _forget(provided):
    """Other tests import the REAL reference modules under their own stand-ins (oracle/ref_harness.py): leave no
    `src.*` module, alias or stand-in of this file behind."""
    for k in list(sys.modules):
        if k == "src" or k.startswith("src.") or k in provided or k in ("decord", "av", "lmdb", "spacy") and \
                not hasattr(sys.modules[k], "__file__"):
            del sys.modules[k]


@pytest.fixture
def reference_scripts():
    if not os.path.isdir(REF + "/src"):
        pytest.skip("/root/reference not present")
    saved = _set_aside()
    _fake_data_libs()
    provided = shims.install(alias_models=True, optional_stubs=True, force=True)
    sys.path.insert(0, REF)
    try:
        yield provided
    finally:
        sys.path.remove(REF)
        _forget(provided)
        sys.modules.update(saved)


def _write_cfgs(tmp_path, T, img):
    """Model configs in the reference's JSON schema, shrunk through the optional keys alpro_b200 understands."""
    cfg, b, v = _tiny_models_cfg(T=T, img=img)
    mpath, vpath = str(tmp_path / "base_model.json"), str(tmp_path / "timesformer.json")
    json.dump(b, open(mpath, "w"))
    json.dump({k: v[k] for k in v if k not in ("num_frm", "img_size")}, open(vpath, "w"))
    return cfg, mpath, vpath


@pytest.mark.reference
def test_unchanged_retrieval_script_builds_our_model(reference_scripts, tmp_path):
    import importlib
    from easydict import EasyDict as edict
    rv = importlib.import_module("src.tasks.run_video_retrieval")           # the reference file, unmodified
    assert rv.AlproForVideoTextRetrieval is modeling.AlproForVideoTextRetrieval
    assert getattr(rv.hvd, "__name__", "").endswith("shims.hvd") and rv.amp.__name__.endswith("shims.amp")
    cfg, mpath, vpath = _write_cfgs(tmp_path, T=2, img=64)
    # an e2e checkpoint with another grid / frame count: exercises the reference's own
    # load_state_dict_with_pos_embed_resizing (src/utils/load_save.py:73-140) on OUR module tree
    src_cfg, sb, sv = _tiny_models_cfg(T=4, img=48)
    donor = modeling.AlproForVideoTextRetrieval(sb, sv)
    wpath = str(tmp_path / "e2e.pt")
    torch.save(donor.state_dict(), wpath)
    rcfg = edict(model_config=mpath, visual_model_cfg=vpath, num_frm=2, crop_img_size=64, img_input_format="RGB",
                 e2e_weights_path=wpath, visual_weights_path=None, bert_weights_path=None)
    model = rv.setup_model(rcfg, device="cpu")
    assert isinstance(model, modeling.AlproForVideoTextRetrieval)
    dsd = donor.state_dict()
    assert torch.equal(model.text_encoder.bert.encoder.layer[1].output.dense.weight,
                       dsd["text_encoder.bert.encoder.layer.1.output.dense.weight"])
    assert torch.equal(model.visual_encoder.model.pos_embed,
                       modeling.resize_spatial_embedding(dsd["visual_encoder.model.pos_embed"], 16))
    assert torch.equal(model.visual_encoder.model.time_embed,
                       modeling.resize_temporal_embedding(dsd["visual_encoder.model.time_embed"], 2))
    # forward_step reaches our forward: on a CPU-only container that is the loud no-fallback error
    batch = synth.synth_batch("retrieval", 2, 2, 64, 8, cfg["bert"]["vocab_size"], seed=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        rv.forward_step(model, batch)
    # the trainer's optimizer plumbing runs on our parameter names (src/optimization/utils.py)
    from src.optimization.utils import setup_e2e_optimizer
    ocfg = edict(optim="adamw", learning_rate=1e-4, weight_decay=1e-3, betas=[0.9, 0.98], cnn_learning_rate=1e-4,
                 cnn_weight_decay=1e-3, cnn_sgd_momentum=0.9, cnn_lr_decay="linear", transformer_lr_mul=1.0,
                 transformer_lr_mul_prefix="")
    try:
        opt = setup_e2e_optimizer(model, ocfg)
    except (AttributeError, KeyError) as e:
        pytest.skip(f"optimizer config schema differs: {e}")
    dopt = rv.hvd.DistributedOptimizer(opt, named_parameters=model.named_parameters(),
                                       compression=rv.hvd.Compression.none)
    rv.hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    model2, dopt = rv.amp.initialize(model, dopt, enabled=False, opt_level="O2", keep_batchnorm_fp32=True)
    assert model2 is model


@pytest.mark.reference
def test_unchanged_pretrain_script_builds_our_model(reference_scripts, tmp_path):
    import importlib
    from easydict import EasyDict as edict
    rp = importlib.import_module("src.pretrain.run_pretrain_sparse")
    assert rp.AlproForPretrain is modeling.AlproForPretrain
    cfg, mpath, vpath = _write_cfgs(tmp_path, T=2, img=64)
    _, b, v = _tiny_models_cfg(T=2, img=64)
    teacher = modeling.Prompter(b, v)
    tpath = str(tmp_path / "teacher.pt")
    torch.save(teacher.state_dict(), tpath)
    vis_spec = synth.visual_spec("", cfg["vis"]["d"], cfg["vis"]["depth"], 8, 196, 16)    # K600-style 8 x 224^2
    vw = str(tmp_path / "vis.pyth")
    torch.save({"model_state": {"model." + k: synth.synth_tensor(k, shp, 4) for k, shp in vis_spec.items()}}, vw)
    rcfg = edict(model_config=mpath, visual_model_cfg=vpath, num_frm=2, crop_img_size=64, img_input_format="RGB",
                 model_type="pretrain", max_n_example_per_group=1, num_entities=cfg["num_entities"],
                 e2e_weights_path=None, visual_weights_path=vw, teacher_weights_path=tpath, use_itm=True)
    model = rp.setup_model(rcfg, device="cpu")
    assert isinstance(model, modeling.AlproForPretrain) and isinstance(model.prompter, modeling.Prompter)
    assert torch.equal(model.prompter.itm_head.weight, teacher.itm_head.weight)
    assert model.visual_encoder.model.pos_embed.shape == (1, 17, cfg["vis"]["d"])
    batch = synth.synth_batch("pretrain", 2, 2, 64, 8, cfg["bert"]["vocab_size"], seed=1,
                              num_entities=cfg["num_entities"])
    with pytest.raises(RuntimeError, match="CUDA only"):
        rp.forward_step(rcfg, model, batch)


def test_launcher_runs_a_script_against_our_classes(tmp_path):
    """python -m alpro_b200.launch <script>: the script's reference-style imports resolve to alpro_b200 / the stand-ins."""
    import subprocess
    (tmp_path / "src" / "modeling").mkdir(parents=True)
    (tmp_path / "src" / "__init__.py").write_text("")
    (tmp_path / "src" / "modeling" / "__init__.py").write_text("")
    script = tmp_path / "run_dummy.py"
    script.write_text(
        "import sys\n"
        "import horovod.torch as hvd\n"
        "from apex import amp\n"
        "from src.modeling.alpro_models import AlproForPretrain, AlproForVideoTextRetrieval, Prompter\n"
        "hvd.init()\n"
        "assert __name__ == '__main__' and sys.argv[1:] == ['--config', 'x.json']\n"
        "print('LAUNCH_OK', AlproForPretrain.__module__, hvd.size(), amp.initialize(1, 2, enabled=False))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, "-m", "alpro_b200.launch", str(script), "--config", "x.json"], cwd=str(tmp_path),
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "LAUNCH_OK alpro_b200.modeling 1 (1, 2)" in r.stdout, r.stdout
