"""GPU: the fused hard-negative sampler (alpro_neg_sample; reference: torch.multinomial(weights[b], 1) per row,
alpro_models.py:301-316, 833-844): Philox4x32-10 known-answer vectors, the weights it reports, determinism, the
diagonal exclusion, and a chi-square test of the drawn distribution against the softmax weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox_ref(ctr, key, rounds=10):
    c, k = list(ctr), list(key)
    for _ in range(rounds):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
        k = [(k[0] + W0) & 0xffffffff, (k[1] + W1) & 0xffffffff]
    return c


def test_philox_known_answers():
    """Random123 kat_vectors (philox4x32 10) + a sweep against the Python restatement."""
    from alpro_b200 import ops
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        assert philox_ref(c, k) == want
    g = torch.Generator().manual_seed(0)
    rows = torch.randint(0, 2 ** 32, (64, 6), generator=g, dtype=torch.int64)
    for i, (c, k, _) in enumerate(kat):
        rows[i] = torch.tensor(c + k)
    inp32 = ((rows + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32).to(DEV)   # the 32-bit patterns as int32
    out = torch.empty(64, 4, dtype=torch.int32, device=DEV)
    ops.philox4x32_10(inp32, out)
    got = (out.to(torch.int64) & 0xffffffff).cpu().tolist()
    for i in range(64):
        r = rows[i].tolist()
        assert got[i] == philox_ref(r[:4], r[4:]), i


def test_neg_sample_weights_determinism_and_diagonal():
    from alpro_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(1)
    for b, G, col0 in ((2, 2, 0), (7, 21, 7), (32, 256, 64), (100, 100, 0)):
        sim = torch.randn(b, G, device=DEV, generator=g) * 4
        w = torch.empty(b, b, device=DEV)
        idx = torch.empty(b, dtype=torch.int64, device=DEV)
        ops.neg_sample(sim, col0, b, 1234567891011, 5, idx, w)
        blk = sim[:, col0:col0 + b].clone()
        blk.fill_diagonal_(-float("inf"))
        assert torch.allclose(w, blk.softmax(dim=1), atol=1e-6, rtol=1e-5)
        assert int(idx.min()) >= 0 and int(idx.max()) < b
        assert not bool((idx == torch.arange(b, device=DEV)).any())          # never the positive itself
        idx2 = torch.empty_like(idx)
        ops.neg_sample(sim, col0, b, 1234567891011, 5, idx2)
        assert torch.equal(idx, idx2)                                         # pure function of (seed, draw, row)
    # different draws / seeds decorrelate
    sim = torch.zeros(64, 64, device=DEV)
    a, c, d = (torch.empty(64, dtype=torch.int64, device=DEV) for _ in range(3))
    ops.neg_sample(sim, 0, 64, 99, 0, a)
    ops.neg_sample(sim, 0, 64, 99, 1, c)
    ops.neg_sample(sim, 0, 64, 100, 0, d)
    assert float((a != c).float().mean()) > 0.8 and float((a != d).float().mean()) > 0.8


def test_neg_sample_distribution_chi_square():
    """20 000 draws per row against the softmax weights: chi-square below the 99.9 % quantile for every row."""
    from alpro_b200 import ops
    b, n = 12, 20000
    g = torch.Generator(device=DEV).manual_seed(2)
    sim = torch.randn(b, b, device=DEV, generator=g) * 1.5
    blk = sim.clone()
    blk.fill_diagonal_(-float("inf"))
    wref = blk.softmax(dim=1).double().cpu()
    counts = torch.zeros(b, b, dtype=torch.float64)
    idx = torch.empty(b, dtype=torch.int64, device=DEV)
    draws = torch.empty(n, b, dtype=torch.int64, device=DEV)
    for t in range(n):
        ops.neg_sample(sim, 0, b, 42, t, draws[t])
    torch.cuda.synchronize()
    for r in range(b):
        counts[r] = torch.bincount(draws[:, r].cpu(), minlength=b).double()
    crit = 31.26      # chi-square 99.9 % quantile, 10 degrees of freedom (b - 2: diagonal excluded, counts sum to n)
    for r in range(b):
        keep = [c for c in range(b) if c != r]
        exp = wref[r, keep] * n
        chi = float((((counts[r, keep] - exp) ** 2) / exp).sum())
        assert chi < crit, (r, chi)
        assert counts[r, r] == 0


def test_engine_uses_fused_sampler_by_default():
    from oracle import configs
    from tests import helpers
    from tests.test_gpu_parity import build_cuda_model, to_cuda
    cfg = configs.GOLDEN["tiny_retrieval"]
    spec, sd, batch = helpers.make_inputs(cfg)
    model = build_cuda_model(cfg, sd)
    model.engine.sampler = None
    B = cfg["B"]
    seen = set()
    for _ in range(6):
        out = model(to_cuda(batch))
        nv, nt = out["_neg_video"], out["_neg_text"]
        assert nv.dtype == torch.int64 and nv.is_cuda
        ar = torch.arange(B, device=DEV)
        assert not bool((nv == ar).any()) and not bool((nt == ar).any())
        seen.add(tuple(nv.tolist() + nt.tolist()))
    assert len(seen) > 1          # the draw advances from step to step
    (out["itc_loss"] + out["itm_loss"]).backward()
