"""CPU check of the batched DropPath draw (alpro_b200/engine.py::VisualEncoder._drop_path_all): given the same per-sample
masks it must expand to exactly the row factors of the per-block reference implementation (_drop_path_scales, which
follows the three sample granularities of vit.py:157,181,212 / vit_utils.py:137-162), and a fresh draw must have the
right support and keep rates."""
import torch

from alpro_b200.engine import VisualEncoder


def _engine(depth=4):
    return VisualEncoder("visual_encoder.model.", dict(d=64, depth=depth, heads=1, T=2, img=32, patch=16), torch.float16)


def test_batched_expansion_equals_per_block_reference():
    eng = _engine()
    B, N, T, rate = 3, 4, 2, 0.3
    dev = torch.device("cpu")
    raw, ref = {}, {}
    for i in range(eng.depth):
        torch.manual_seed(100 + i)
        d = eng._drop_path_scales(i, rate, B, N, T, dev)
        ref[i] = d
        raw[i] = None if d is None else (d["raw"]["m_t"], d["raw"]["m_s"], d["raw"]["m_m"])
    got = eng._drop_path_all(rate, B, N, T, dev, raw=raw)
    assert got[0] is None and ref[0] is None          # block 0 has rate 0
    for i in range(1, eng.depth):
        for k in ("rs_t", "rsa", "rsb", "rs_m", "m_s", "cls_w"):
            assert got[i][k].is_contiguous() and got[i][k].dtype == torch.float32
            assert torch.equal(got[i][k], ref[i][k]), (i, k)
        for k in ("m_t", "m_s", "m_m"):
            assert torch.equal(got[i]["raw"][k], ref[i]["raw"][k])


def test_batched_draw_statistics():
    eng = _engine(depth=3)
    B, N, T, rate = 64, 16, 2, 0.4
    torch.manual_seed(0)
    got = eng._drop_path_all(rate, B, N, T, torch.device("cpu"))
    assert got[0] is None
    for i in (1, 2):
        keep = 1.0 - rate * i / 2
        for k in ("m_t", "m_s", "m_m"):
            m = got[i]["raw"][k]
            vals = torch.unique(m)
            assert all(abs(float(v)) < 1e-6 or abs(float(v) - 1.0 / keep) < 1e-5 for v in vals), (i, k, vals)
        frac = float((got[i]["raw"]["m_t"] > 0).float().mean())
        assert abs(frac - keep) < 0.06, (i, frac, keep)
        Sc = 1 + N * T
        assert got[i]["rs_t"].numel() == B * Sc and float(got[i]["rs_t"].view(B, Sc)[:, 0].abs().max()) == 0.0
        assert float(got[i]["rsa"].view(B, Sc)[:, 0].min()) == 1.0
