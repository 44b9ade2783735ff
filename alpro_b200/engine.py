"""Host-side orchestration of the ALPRO forward/backward on the sm_100a kernels (alpro_b200/csrc/*).

This is explicit forward + hand-derived backward (no autograd inside): every tensor below is a device buffer handed to
a C-ABI kernel through alpro_b200.ops. torch is used for allocation, streams and a few index/scalar glue ops only.

Layout / precision choices:
  * TimeSformer tokens live in ONE canonical layout [B*(1+N*T), d] with row = b*(1+N*T) + 1 + n*T + t (the reference's
    'b (h w t) m', vit.py:147); temporal and spatial attention read it through index arithmetic instead of rearranges.
  * residual stream, LayerNorm / softmax statistics and all head math are fp32; GEMM operands are 16-bit (fp16 by
    default) with fp32 TMEM accumulation; the backward pass carries gradients multiplied by a static loss scale S.
"""
import math
import os

import torch

from . import ops
from .ops import ACT_GELU, ACT_GELU_GRAD, KMAJOR, MNMAJOR

EPS_VIT = 1e-6


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class OperandCache:
    """16-bit operand copies of fp32 parameters.

    A copy is refreshed when the parameter's version counter changes OR when the cache epoch has moved. The version
    counter alone is not enough: optimizers that update through `p.data` (the reference's AdamW,
    src/optimization/adamw.py:80-98, EMA / `p.data.copy_()` patterns) leave `_version` untouched. AlproEngine therefore
    bumps the epoch at the start of every gradient-tracking forward and at the end of every backward, so each training
    step re-casts its operands exactly once (~1.4 GB of traffic, ~0.3 ms) and an eval forward after a training step
    never sees stale weights. Only repeated no-grad forwards reuse the copies; after an out-of-band `.data` update
    between two no-grad forwards call `model.invalidate_operands()`."""

    def __init__(self, dtype):
        self.dtype = dtype
        self._c = {}
        self._epoch = 0
        self._plain = {}       # name -> (parameter, 16-bit copy): entries that are a straight cast (batched refresh)
        self._table = None     # (signature, device int64 chunk table, number of chunks)

    def _key(self, ps):
        # frozen parameters (the pretraining teacher) are never touched by an optimizer: version counter only
        return tuple((p.data_ptr(), p._version, self._epoch if p.requires_grad else -1) for p in ps)

    def invalidate(self):
        """Force a refresh of every operand copy (used after an out-of-band parameter update, e.g. FusedAdamW)."""
        self._epoch += 1

    def refresh(self):
        """invalidate() + re-cast every straight-cast operand seen so far with ONE launch (alpro_cast_f32_to_16_multi)
        instead of one per tensor on its next use (~200 launches per training step). Composed operands (W_fc W_proj)
        and tensors that appeared since the table was built refresh lazily as before."""
        self._epoch += 1
        if not self._plain:
            return
        live = {n: (p, d) for n, (p, d) in self._plain.items() if p.requires_grad}
        if not live:
            return
        sig = tuple((n, p.data_ptr(), d.data_ptr(), p.numel()) for n, (p, d) in live.items())
        if self._table is None or self._table[0] != sig:
            rows = []
            for n, (p, d) in live.items():
                src, dst, numel = p.data_ptr(), d.data_ptr(), p.numel()
                if src % 16 or dst % 8 or not p.is_contiguous():
                    self._table = None
                    return                      # unusual layout: leave everything to the lazy path
                for o in range(0, numel, ops.CAST_CHUNK):
                    rows.append((src + 4 * o, dst + 2 * o, min(ops.CAST_CHUNK, numel - o)))
            dev = next(iter(live.values()))[1].device
            self._table = (sig, torch.tensor(rows, dtype=torch.int64).to(dev), len(rows))
        ops.cast16_multi(self._table[1], self._table[2], ops._FMT[self.dtype])
        for n, (p, d) in live.items():          # mark fresh: the next get() is a hit
            ent = self._c.get(n)
            if ent is not None and ent[1] is d:
                self._c[n] = (self._key((p,)), d)
            else:                               # a row slice of a concatenated operand (fused q|k|v)
                self._fresh_parts = getattr(self, "_fresh_parts", {})
                self._fresh_parts[n] = self._epoch

    def get(self, name, p, shape2d=None):
        ent = self._c.get(name)
        key = self._key((p,))
        if ent is None or ent[0] != key:
            src = p.detach()
            dst = ent[1] if ent is not None else _empty(shape2d or tuple(src.shape), self.dtype, src.device)
            ops.cast16(src.contiguous().view(-1), dst.view(-1))
            ent = (key, dst)
            self._c[name] = ent
            if src.is_contiguous():
                self._plain[name] = (p, dst)
        return ent[1]

    def get_fused_linear(self, name, w2, w1, b1):
        """Operand of the composition x -> W2 (W1 x + b1): (W2 W1 as a 16-bit [out, in] operand, W2 b1 as fp32).
        TimeSformer applies temporal_fc directly to the (DropPath-scaled) output of temporal_attn.proj (vit.py:157-161),
        so one GEMM with the composed weight stands for the two Linear maps. Refreshed when a parameter changes."""
        ent = self._c.get(name)
        key = self._key((w2, w1, b1))
        if ent is None or ent[0] != key:
            w2_16 = self.get(name + "#w2", w2)
            w1_16 = self.get(name + "#w1", w1)
            n_out, n_in = w2.shape[0], w1.shape[1]
            wc = ent[1][0] if ent is not None else _empty((n_out, n_in), self.dtype, w2.device)
            c1 = ent[1][1] if ent is not None else _empty((n_out,), torch.float32, w2.device)
            # Wc[i, j] = sum_k W2[i, k] W1[k, j]: A = W2 (K-major), B = W1 stored [k][j] = MN-major
            ops.gemm16(w2_16, w1_16, b_layout=MNMAJOR, out16=wc)
            # c1[i] = sum_k b1[k] W2[i, k]  (fp32)
            ops.small_linear_fwd(b1.detach().view(1, -1), b1.numel(), w2.detach(), None, c1.view(1, -1), 1, n_out,
                                 w2.shape[1])
            ent = (key, (wc, c1))
            self._c[name] = ent
        return ent[1]

    def get_cat(self, name, ps):
        """Row-concatenated operand (fused BERT q|k|v weight) and fp32 bias concatenation."""
        ent = self._c.get(name)
        key = self._key(ps)
        if ent is None or ent[0] != key:
            rows = sum(p.shape[0] for p in ps)
            if ps[0].dim() == 2:
                dst = ent[1] if ent is not None else _empty((rows, ps[0].shape[1]), self.dtype, ps[0].device)
                fresh = getattr(self, "_fresh_parts", {})
                r = 0
                for i, p in enumerate(ps):
                    part = dst[r:r + p.shape[0]]
                    pname = f"{name}#{i}"
                    reg = self._plain.get(pname)
                    done = (reg is not None and reg[0] is p and reg[1].data_ptr() == part.data_ptr()
                            and fresh.get(pname) == self._epoch)          # already re-cast by refresh() this epoch
                    if not done:
                        ops.cast16(p.detach().contiguous().view(-1), part.view(-1))
                    if p.is_contiguous():
                        self._plain[pname] = (p, part)
                    r += p.shape[0]
            else:
                dst = torch.cat([p.detach() for p in ps]).contiguous()
            ent = (key, dst)
            self._c[name] = ent
        return ent[1]


class GradStore:
    """One flat zero-initialised fp32 buffer with a view per parameter (single memset per backward).
    `groups` lists parameters that must be adjacent (BERT q|k|v) so that one fused wgrad GEMM can fill all three."""

    def __init__(self, named_params, device, groups=(), alloc=None):
        """alloc(numel, device) -> zeroed flat fp32 tensor (data-parallel runs hand out symmetric memory that the
        peer-memory gradient reducer can address on every rank); default: torch.zeros."""
        shapes = {n: tuple(p.shape) for n, p in named_params}
        order = []
        seen = set()
        self.groups = {}
        for gname, names in groups:
            if all(n in shapes for n in names):
                order.extend(names)
                seen.update(names)
                self.groups[gname] = names
        # Region order = the order in which the backward pass FINISHES gradients, so that a data-parallel all-reduce of
        # one contiguous slice can start while later regions are still being computed (comm.BucketedAllReduce):
        #   region 0: everything outside the visual encoder (heads, fusion/text BERT, embeddings)
        #   region 1+k: visual block (depth-1-k) [+ the final norm for k = 0]; last region: patch/pos/time/cls embeddings
        rest = [n for n in shapes if n not in seen]
        vis = [n for n in rest if n.startswith("visual_encoder.")]
        nonvis = [n for n in rest if not n.startswith("visual_encoder.")]

        def vis_key(n):
            if ".blocks." in n:
                return (1, -int(n.split(".blocks.")[1].split(".")[0]))
            if ".model.norm." in n or ".model.head." in n:
                return (0, 0)
            return (2, 0)
        vis.sort(key=vis_key)                      # stable: keeps the module order inside a block
        order.extend(nonvis)
        n_nonvis = len(order)
        order.extend(vis)
        total = 0
        self.offsets = {}
        self.regions = []                          # [(lo, hi, tag)] element ranges in `flat`, in completion order
        cur_tag, cur_lo = ("nonvis",), 0
        for idx, n in enumerate(order):
            tag = ("nonvis",) if idx < n_nonvis else vis_key(n)
            if tag != cur_tag:
                self.regions.append((cur_lo, total, cur_tag))
                cur_tag, cur_lo = tag, total
            numel = 1
            for v in shapes[n]:
                numel *= v
            assert numel % 4 == 0 or n not in seen, "grouped parameters must keep 16-byte alignment"
            self.offsets[n] = (total, numel, shapes[n])
            total += numel if n in seen else (numel + 3) // 4 * 4
        self.regions.append((cur_lo, total, cur_tag))
        self.flat = alloc(total, device) if alloc is not None else torch.zeros(total, dtype=torch.float32, device=device)

    def region_end(self, tag):
        for lo, hi, t in self.regions:
            if t == tag:
                return hi
        return None

    def __getitem__(self, n):
        o, numel, shape = self.offsets[n]
        return self.flat[o:o + numel].view(shape)

    def group(self, gname):
        """Contiguous [sum(rows), ...] view over a fused group."""
        names = self.groups[gname]
        o = self.offsets[names[0]][0]
        numel = sum(self.offsets[n][1] for n in names)
        shape0 = self.offsets[names[0]][2]
        rows = sum(self.offsets[n][2][0] for n in names)
        return self.flat[o:o + numel].view((rows,) + tuple(shape0[1:]))

    def __contains__(self, n):
        return n in self.offsets


def bert_grad_groups(prefix, cfg):
    groups = []
    for i in range(cfg["num_hidden_layers"]):
        l = f"{prefix}bert.encoder.layer.{i}."
        groups.append((l + "qkv.w", [l + f"attention.self.{n}.weight" for n in ("query", "key", "value")]))
        groups.append((l + "qkv.b", [l + f"attention.self.{n}.bias" for n in ("query", "key", "value")]))
    return groups


# =====================================================================================================================
class VisualEncoder:
    """TimeSformer divided space-time encoder (vit.py:242-382, 419-503) on the canonical token layout."""

    def __init__(self, prefix, vis, dtype):
        self.p = prefix  # e.g. 'visual_encoder.model.'
        self.d, self.depth, self.heads = vis["d"], vis["depth"], vis["heads"]
        self.patch = vis["patch"]
        self.dtype = dtype
        # ImageNorm constants for uint8 inputs (config img_pixel_mean/std, config_release/msrvtt_ret.json:19-20)
        self.img_mean = vis.get("img_mean", (0.48145466, 0.4578275, 0.40821073))
        self.img_std = vis.get("img_std", (0.26862954, 0.26130258, 0.27577711))
        assert self.d == self.heads * 64, "kernels assume head_dim 64"

    # ---- parameter helpers
    def _pos_time(self, P, N, T):
        """pos/time embeddings, nearest-resized when the grid / frame count differ (vit.py:328-355)."""
        pos = P[self.p + "pos_embed"].detach()[0]
        tim = P[self.p + "time_embed"].detach()[0]
        pos_idx = tim_idx = None
        if pos.shape[0] != N + 1:
            Pg = int(round(math.sqrt(pos.shape[0] - 1)))
            g = int(round(math.sqrt(N)))
            src = (torch.arange(g, device=pos.device).float() * (Pg / g)).floor().long().clamp_(max=Pg - 1)
            grid = (src[:, None] * Pg + src[None, :]).reshape(-1) + 1
            pos_idx = torch.cat([torch.zeros(1, dtype=torch.long, device=pos.device), grid])
            pos = pos[pos_idx].contiguous()
        if tim.shape[0] != T:
            tim_idx = (torch.arange(T, device=tim.device).float() * (tim.shape[0] / T)).floor().long()
            tim = tim[tim_idx].contiguous()
        return pos.contiguous(), tim.contiguous(), pos_idx, tim_idx

    def _drop_path_scales(self, i, rate, B, N, T, dev):
        """Per-sample stochastic-depth factors mask/keep (DropPath, vit_utils.py:137-162; rate linearly scaled over
        depth, vit.py:272-277) expanded to token rows. The three branches use different sample granularity exactly as
        the reference's mask shape (x.shape[0],1,1): temporal per (b,h,w), spatial per (b,t), MLP per b."""
        p = rate * i / max(self.depth - 1, 1)
        if p <= 0.0:
            return None
        keep = 1.0 - p
        Sc = 1 + N * T
        m_t = torch.bernoulli(torch.full((B, N), keep, device=dev)) / keep
        m_s = torch.bernoulli(torch.full((B, T), keep, device=dev)) / keep
        m_m = torch.bernoulli(torch.full((B,), keep, device=dev)) / keep
        zero = torch.zeros(B, 1, device=dev)
        one = torch.ones(B, 1, device=dev)
        rs_t = torch.cat([zero, m_t.view(B, N, 1).expand(B, N, T).reshape(B, N * T)], 1).reshape(-1).contiguous()
        sp = m_s.view(B, 1, T).expand(B, N, T).reshape(B, N * T)
        rsa = torch.cat([one, sp], 1).reshape(-1).contiguous()
        rsb = torch.cat([m_s.mean(dim=1, keepdim=True), sp], 1).reshape(-1).contiguous()
        rs_m = m_m.view(B, 1).expand(B, Sc).reshape(-1).contiguous()
        return dict(rs_t=rs_t, rsa=rsa, rsb=rsb, rs_m=rs_m, m_s=m_s.reshape(-1).contiguous(),
                    cls_w=(m_s / T).reshape(-1).contiguous(), raw=dict(m_t=m_t, m_s=m_s, m_m=m_m))

    def _drop_path_all(self, rate, B, N, T, dev, raw=None):
        """The stochastic-depth factors of ALL blocks from one batched draw: the per-block version above costs ~13 tiny
        launches per block (bernoulli / div / cat / copies), ~140 per training step; this one ~20 per step. Returns a list
        with one dict (same keys as _drop_path_scales) or None per block. `raw` (tests): per-block (m_t, m_s, m_m) masks
        already divided by keep, or None for blocks without DropPath, instead of a fresh draw."""
        depth = self.depth
        ps = [rate * i / max(depth - 1, 1) for i in range(depth)]
        active = [i for i in range(depth) if ps[i] > 0.0]
        out = [None] * depth
        if not active:
            return out
        A, Sc = len(active), 1 + N * T
        if raw is None:
            ck = (rate, depth, str(dev))
            if getattr(self, "_dp_keep_key", None) != ck:   # per-block keep probabilities, uploaded once
                self._dp_keep = torch.tensor([1.0 - ps[i] for i in active], device=dev).view(A, 1, 1)
                self._dp_keep_key = ck
            keep = self._dp_keep
            m = (torch.rand(A, B, N + T + 1, device=dev) < keep).float() / keep
            m_t, m_s, m_m = m[:, :, :N], m[:, :, N:N + T], m[:, :, N + T]
        else:
            m_t = torch.stack([raw[i][0] for i in active])
            m_s = torch.stack([raw[i][1] for i in active])
            m_m = torch.stack([raw[i][2] for i in active])
        rs_t = torch.zeros(A, B, Sc, device=dev)
        rs_t[:, :, 1:] = m_t.unsqueeze(-1).expand(A, B, N, T).reshape(A, B, N * T)
        sp = m_s.unsqueeze(2).expand(A, B, N, T).reshape(A, B, N * T)
        rsa = torch.ones(A, B, Sc, device=dev)
        rsa[:, :, 1:] = sp
        rsb = torch.empty(A, B, Sc, device=dev)
        rsb[:, :, 0] = m_s.mean(dim=2)
        rsb[:, :, 1:] = sp
        rs_m = m_m.unsqueeze(-1).expand(A, B, Sc).contiguous()
        m_s_c = m_s.contiguous()
        cls_w = m_s_c / T
        for a, i in enumerate(active):
            out[i] = dict(rs_t=rs_t[a].reshape(-1), rsa=rsa[a].reshape(-1), rsb=rsb[a].reshape(-1),
                          rs_m=rs_m[a].reshape(-1), m_s=m_s_c[a].reshape(-1), cls_w=cls_w[a].reshape(-1),
                          raw=dict(m_t=m_t[a], m_s=m_s[a], m_m=m_m[a]))
        return out

    def forward(self, P, W, frames, save, drop_path_rate=0.0):
        """frames fp32 [B,T,3,H,W] -> video_embeds fp32 [B, 1+N, d]; ctx holds what the backward needs.
        drop_path_rate > 0 enables train-mode stochastic depth."""
        B, T, C, H, Wd = frames.shape
        dev = frames.device
        d, heads, dt = self.d, self.heads, self.dtype
        N = (H // self.patch) * (Wd // self.patch)
        Sc = 1 + N * T
        M = B * Sc
        p = self.p
        scale = 64 ** -0.5
        ctx = {"B": B, "T": T, "N": N, "blocks": []} if save else None

        frames = frames.contiguous()
        patches = _empty((M, 3 * self.patch * self.patch), dt, dev)
        if frames.dtype == torch.uint8:   # raw frames: ImageNorm fused into the patch gather
            ops.patchify_u8(frames, patches, self.patch, self.img_mean, self.img_std)
        else:
            ops.patchify(frames, patches, self.patch)
        Wpe = W.get(p + "patch_embed.proj.weight", P[p + "patch_embed.proj.weight"], (d, 3 * self.patch ** 2))
        proj = _empty((M, d), torch.float32, dev)
        ops.gemm16(patches, Wpe, bias=P[p + "patch_embed.proj.bias"].detach(), out32=proj)
        pos, tim, pos_idx, tim_idx = self._pos_time(P, N, T)
        x = _empty((M, d), torch.float32, dev)
        ops.vit_embed_fwd(proj, P[p + "cls_token"].detach().view(-1), pos, tim, x, B, N, T, d)
        del proj
        if save:
            ctx.update(patches=patches, pos_idx=pos_idx, tim_idx=tim_idx)

        # scratch reused across blocks when nothing has to be kept
        scratch = {}

        def buf(name, shape, dtype):
            if save:
                return _empty(shape, dtype, dev)
            t = scratch.get(name)
            if t is None:
                t = scratch[name] = _empty(shape, dtype, dev)
            return t

        # temporal_attn.proj and temporal_fc run as one composed Linear (one 385 MB HBM-bound GEMM less per block in the
        # forward, one dgrad + one wgrad less in the backward); ALPRO_FUSE_TFC=0 keeps the two Linear maps (read per call)
        fuse_tfc = os.environ.get("ALPRO_FUSE_TFC", "1") != "0"
        if save:
            ctx["fuse_tfc"] = fuse_tfc
        dps = self._drop_path_all(drop_path_rate, B, N, T, dev) if drop_path_rate > 0 else [None] * self.depth
        for i in range(self.depth):
            b = f"{p}blocks.{i}."
            g = lambda n: P[b + n].detach()
            w = lambda n: W.get(b + n, P[b + n])
            dp = dps[i]
            # ---- temporal attention branch (vit.py:146-162)
            a_t = buf("a_t", (M, d), dt)
            st_t = buf("st_t", (2, M), torch.float32)
            ops.layernorm_fwd(x, g("temporal_norm1.weight"), g("temporal_norm1.bias"), EPS_VIT, out16=a_t,
                              mean=st_t[0], rstd=st_t[1])
            qkv_t = buf("qkv_t", (M, 3 * d), dt)
            ops.gemm16(a_t, w("temporal_attn.qkv.weight"), bias=g("temporal_attn.qkv.bias"), out16=qkv_t)
            o_t = buf("o_t", (M, d), dt)
            ops.temporal_attn_fwd(qkv_t, o_t, B, N, T, heads, scale)
            x1 = buf("x1", (M, d), torch.float32)
            if fuse_tfc:
                # proj -> DropPath -> temporal_fc as ONE GEMM: x1 = x + s (o Wc^T + W_fc b_proj) + b_fc, Wc = W_fc W_proj
                p_t = None
                wc, c1 = W.get_fused_linear(b + "temporal_fc*proj", P[b + "temporal_fc.weight"],
                                            P[b + "temporal_attn.proj.weight"], P[b + "temporal_attn.proj.bias"])
                ops.gemm16(o_t, wc, bias=c1, bias2=g("temporal_fc.bias"), resid=x, skip_period=Sc, out32=x1,
                           row_scale=dp["rs_t"] if dp else None)
            else:
                p_t = buf("p_t", (M, d), dt)
                ops.gemm16(o_t, w("temporal_attn.proj.weight"), bias=g("temporal_attn.proj.bias"), out16=p_t,
                           row_scale=dp["rs_t"] if dp else None)
                ops.gemm16(p_t, w("temporal_fc.weight"), bias=g("temporal_fc.bias"), resid=x, skip_period=Sc, out32=x1)
            # ---- spatial attention branch (vit.py:165-196)
            a_s = buf("a_s", (M, d), dt)
            st_s = buf("st_s", (2, M), torch.float32)
            ops.layernorm_fwd(x1, g("norm1.weight"), g("norm1.bias"), EPS_VIT, out16=a_s, mean=st_s[0], rstd=st_s[1])
            qkv_s = buf("qkv_s", (M, 3 * d), dt)
            ops.gemm16(a_s, w("attn.qkv.weight"), bias=g("attn.qkv.bias"), out16=qkv_s)
            o_s = buf("o_s", (M, d), dt)
            cls_o = buf("cls_o", (B * T, d), dt)
            lse = buf("lse", (B * T, heads, 1 + N), torch.float32)
            ops.seq_attn_fwd(qkv_s, None, o_s, cls_o, lse, 1 + N, B * T, heads, T, T, Sc, scale)
            ops.cls_mean_fwd(cls_o, o_s, B, T, Sc, d, dp["m_s"] if dp else None)
            x2 = buf("x2", (M, d), torch.float32)
            ops.gemm16(o_s, w("attn.proj.weight"), bias=g("attn.proj.bias"), resid=x1, out32=x2,
                       row_scale=dp["rsa"] if dp else None, row_scale_bias=dp["rsb"] if dp else None)
            # ---- MLP (vit.py:198-212)
            a_m = buf("a_m", (M, d), dt)
            st_m = buf("st_m", (2, M), torch.float32)
            ops.layernorm_fwd(x2, g("norm2.weight"), g("norm2.bias"), EPS_VIT, out16=a_m, mean=st_m[0], rstd=st_m[1])
            hdn = buf("hdn", (M, 4 * d), dt)
            pre = buf("pre", (M, 4 * d), dt) if save else None
            ops.gemm16(a_m, w("mlp.fc1.weight"), bias=g("mlp.fc1.bias"), act=ACT_GELU, out16=hdn, out16b=pre)
            x3 = buf("x3", (M, d), torch.float32) if save else x  # inference: write back into x
            ops.gemm16(hdn, w("mlp.fc2.weight"), bias=g("mlp.fc2.bias"), resid=x2, out32=x3,
                       row_scale=dp["rs_m"] if dp else None)
            if save:
                ctx["blocks"].append(dict(dp=dp, x=x, a_t=a_t, st_t=st_t, qkv_t=qkv_t, o_t=o_t, p_t=p_t, x1=x1, a_s=a_s,
                                          st_s=st_s, qkv_s=qkv_s, o_s=o_s, cls_o=cls_o, lse=lse, x2=x2, a_m=a_m, st_m=st_m,
                                          hdn=hdn,
                                          pre=pre))
            x = x3
        xn = _empty((M, d), torch.float32, dev)
        st_f = _empty((2, M), torch.float32, dev)
        ops.layernorm_fwd(x, P[p + "norm.weight"].detach(), P[p + "norm.bias"].detach(), EPS_VIT, out32=xn,
                          mean=st_f[0], rstd=st_f[1])
        ve = _empty((B, 1 + N, d), torch.float32, dev)
        ops.temporal_pool_fwd(xn, ve, B, N, T, d)
        if save:
            ctx.update(x_final=x, st_f=st_f)
        return ve, ctx

    def backward(self, P, W, ctx, d_ve, G, S, hook=None):
        """d_ve: fp32 [B,1+N,d] gradient of video_embeds (already multiplied by the loss scale S). Fills G[...].
        hook(G, end): called whenever the flat gradient prefix [0, end) has become final (after each block)."""
        B, T, N = ctx["B"], ctx["T"], ctx["N"]
        d, heads, dt = self.d, self.heads, self.dtype
        dev = d_ve.device
        Sc = 1 + N * T
        M = B * Sc
        p = self.p
        inv = 1.0 / S
        scale = 64 ** -0.5

        dxn = _empty((M, d), torch.float32, dev)
        ops.temporal_pool_bwd(d_ve.contiguous(), dxn, B, N, T, d)
        dx = _empty((M, d), torch.float32, dev)
        dx16 = _empty((M, d), dt, dev)
        last = f"{p}blocks.{self.depth - 1}."
        dpl = ctx["blocks"][self.depth - 1]["dp"]
        ops.layernorm_bwd(dxn, ctx["x_final"], ctx["st_f"][0], ctx["st_f"][1], P[p + "norm.weight"].detach(), dx, 0,
                          dx16=dx16, dgamma=G[p + "norm.weight"], dbeta=G[p + "norm.bias"], param_scale=inv,
                          colsum=G[last + "mlp.fc2.bias"], dx16_row_scale=dpl["rs_m"] if dpl else None)
        del dxn
        d4 = _empty((M, 4 * d), dt, dev)
        d3 = _empty((M, 3 * d), dt, dev)
        da = _empty((M, d), dt, dev)
        db_ = _empty((M, d), dt, dev)
        scratch = _empty((B * T, 3 * d), torch.float32, dev)
        ones_rows = None
        if ctx.get("fuse_tfc"):
            # per-block accumulators of the composed-weight gradient, zeroed by two fills instead of two per block
            dwc32_all = torch.zeros(self.depth, d, d, device=dev)
            v32_all = torch.zeros(self.depth, d, device=dev)
            dwc16 = _empty((d, d), dt, dev)
            ws = 256.0

        def wgrad(dy16, x16, wname, bname=None, zero_period=0):
            # bname=None: the bias gradient was already produced by the LayerNorm backward that emitted dy16
            ops.gemm16(dy16, x16, a_layout=MNMAJOR, b_layout=MNMAJOR, out32=G[wname].view(G[wname].shape[0], -1),
                       split_k=-1, alpha=inv)
            if bname is not None:
                ops.colsum(dy16, G[bname], inv, zero_period)

        for i in reversed(range(self.depth)):
            b = f"{p}blocks.{i}."
            c = ctx["blocks"][i]
            dp = c["dp"]
            dp_prev = ctx["blocks"][i - 1]["dp"] if i > 0 else None
            g = lambda n: P[b + n].detach()
            w = lambda n: W.get(b + n, P[b + n])
            # ---- MLP (dx16 already carries this block's MLP stochastic-depth factor)
            ops.gemm16(dx16, w("mlp.fc2.weight"), b_layout=MNMAJOR, act=ACT_GELU_GRAD, aux=c["pre"], out16=d4)
            wgrad(dx16, c["hdn"], b + "mlp.fc2.weight")
            ops.gemm16(d4, w("mlp.fc1.weight"), b_layout=MNMAJOR, out16=da)
            wgrad(d4, c["a_m"], b + "mlp.fc1.weight", b + "mlp.fc1.bias")
            ops.layernorm_bwd(da, c["x2"], c["st_m"][0], c["st_m"][1], g("norm2.weight"), dx, 1, dx16=dx16,
                              dgamma=G[b + "norm2.weight"], dbeta=G[b + "norm2.bias"], param_scale=inv,
                              colsum=G[b + "attn.proj.bias"], dx16_row_scale=dp["rsa"] if dp else None,
                              colsum_row_scale=dp["rsb"] if dp else None)
            # ---- spatial attention
            ops.gemm16(dx16, w("attn.proj.weight"), b_layout=MNMAJOR, out16=da)          # d o_s
            wgrad(dx16, c["o_s"], b + "attn.proj.weight")
            ops.seq_attn_bwd(c["qkv_s"], None, c["lse"], c["o_s"], c["cls_o"], da, d3, scratch, 1 + N, B * T, heads, T, T,
                             Sc, scale, cls_weight=dp["cls_w"] if dp else None)
            ops.gemm16(d3, w("attn.qkv.weight"), b_layout=MNMAJOR, out16=da)
            wgrad(d3, c["a_s"], b + "attn.qkv.weight", b + "attn.qkv.bias")
            # dx16 <- grad wrt x1 with cls rows zeroed (the temporal branch never touches cls rows)
            if ctx.get("fuse_tfc"):
                # composed Linear (see forward): dx16 carries the DropPath factor s of the branch, the column sums stay
                # unscaled (they are the gradient of temporal_fc.bias, which sits behind the DropPath)
                if dp and ones_rows is None:
                    ones_rows = torch.ones(M, device=dev)
                ops.layernorm_bwd(da, c["x1"], c["st_s"][0], c["st_s"][1], g("norm1.weight"), dx, 1, dx16=dx16,
                                  zero_period=Sc, dgamma=G[b + "norm1.weight"], dbeta=G[b + "norm1.bias"],
                                  param_scale=inv, colsum=G[b + "temporal_fc.bias"], colsum_zero_period=Sc,
                                  dx16_row_scale=dp["rs_t"] if dp else None, colsum_row_scale=ones_rows if dp else None)
                wc, _ = W.get_fused_linear(b + "temporal_fc*proj", P[b + "temporal_fc.weight"],
                                           P[b + "temporal_attn.proj.weight"], P[b + "temporal_attn.proj.bias"])
                ops.gemm16(dx16, wc, b_layout=MNMAJOR, out16=db_)                             # d o_t
                # dWc = (s * d_res)^T o_t, kept at ws x its true scale so that its 16-bit copy neither overflows (a weight
                # gradient sums over all tokens: the loss scale S of the activation gradients would be too much) nor
                # underflows; v = colsum(s * d_res) stays fp32 at the loss scale
                dwc32, v32 = dwc32_all[i], v32_all[i]
                ops.gemm16(dx16, c["o_t"], a_layout=MNMAJOR, b_layout=MNMAJOR, out32=dwc32, split_k=-1, alpha=ws * inv)
                ops.colsum(dx16, v32, 1.0, 0)
                ops.cast16(dwc32.view(-1), dwc16.view(-1))
                wfc16, wp16 = w("temporal_fc.weight"), w("temporal_attn.proj.weight")
                # dW_fc += dWc W_proj^T + v (x) b_proj ;  dW_proj += W_fc^T dWc ;  db_proj += W_fc^T v
                ops.gemm16(dwc16, wp16, out32=G[b + "temporal_fc.weight"], split_k=2, alpha=1.0 / ws)
                ops.gemm16(wfc16, dwc16, a_layout=MNMAJOR, b_layout=MNMAJOR, out32=G[b + "temporal_attn.proj.weight"],
                           split_k=2, alpha=1.0 / ws)
                ops.small_linear_bwd(v32.view(1, -1), d, None, g("temporal_attn.proj.bias").view(1, -1), d,
                                     P[b + "temporal_fc.weight"].detach(), G[b + "temporal_attn.proj.bias"].view(1, -1),
                                     d, True, G[b + "temporal_fc.weight"], None, True, 1, d, d, alpha=inv)
            else:
                ops.layernorm_bwd(da, c["x1"], c["st_s"][0], c["st_s"][1], g("norm1.weight"), dx, 1, dx16=dx16,
                                  zero_period=Sc, dgamma=G[b + "norm1.weight"], dbeta=G[b + "norm1.bias"],
                                  param_scale=inv, colsum=G[b + "temporal_fc.bias"], colsum_zero_period=Sc)
                # ---- temporal attention
                ops.gemm16(dx16, w("temporal_fc.weight"), b_layout=MNMAJOR, out16=da,        # d p_t
                           row_scale=dp["rs_t"] if dp else None)
                wgrad(dx16, c["p_t"], b + "temporal_fc.weight")
                ops.gemm16(da, w("temporal_attn.proj.weight"), b_layout=MNMAJOR, out16=db_)  # d o_t
                wgrad(da, c["o_t"], b + "temporal_attn.proj.weight", b + "temporal_attn.proj.bias")
            ops.temporal_attn_bwd(c["qkv_t"], db_, d3, B, N, T, heads, scale)
            ops.gemm16(d3, w("temporal_attn.qkv.weight"), b_layout=MNMAJOR, out16=da)
            wgrad(d3, c["a_t"], b + "temporal_attn.qkv.weight", b + "temporal_attn.qkv.bias")
            nxt = G[f"{p}blocks.{i - 1}.mlp.fc2.bias"] if i > 0 else G[p + "patch_embed.proj.bias"]
            ops.layernorm_bwd(da, c["x"], c["st_t"][0], c["st_t"][1], g("temporal_norm1.weight"), dx, 1, dx16=dx16,
                              dgamma=G[b + "temporal_norm1.weight"], dbeta=G[b + "temporal_norm1.bias"],
                              param_scale=inv, colsum=nxt, colsum_zero_period=0 if i > 0 else Sc,
                              dx16_row_scale=dp_prev["rs_m"] if dp_prev else None)
            ctx["blocks"][i] = None  # release saved activations
            if hook is not None and i > 0:
                # every gradient of block i is final here (its mlp.fc2.bias came from the LayerNorm backward that ran
                # before this block's GEMMs); block i-1's region is still open
                hook(G, G.region_end((1, -i)))
        # ---- embeddings (vit.py:324-361) and patch projection
        pos_idx, tim_idx = ctx["pos_idx"], ctx["tim_idx"]
        gpos, gtim = G[p + "pos_embed"][0], G[p + "time_embed"][0]
        dpos = gpos if pos_idx is None else torch.zeros(N + 1, d, device=dev)
        dtim = gtim if tim_idx is None else torch.zeros(T, d, device=dev)
        ops.vit_embed_bwd(dx, G[p + "cls_token"].view(-1), dpos, dtim, B, N, T, d, inv)
        if pos_idx is not None:
            gpos.index_add_(0, pos_idx, dpos)
        if tim_idx is not None:
            gtim.index_add_(0, tim_idx, dtim)
        ops.gemm16(dx16, ctx["patches"], a_layout=MNMAJOR, b_layout=MNMAJOR,
                   out32=G[p + "patch_embed.proj.weight"].view(d, -1), split_k=-1, alpha=inv)


# =====================================================================================================================
class BertEncoder:
    """BERT layers in 'text' / 'fusion' mode (xbert.py:441-630, 832-1081), post-LN, additive key mask."""

    def __init__(self, prefix, cfg, dtype):
        self.p = prefix  # 'text_encoder.'
        self.cfg = cfg
        self.h = cfg["hidden_size"]
        self.heads = cfg["num_attention_heads"]
        self.eps = cfg["layer_norm_eps"]
        self.dtype = dtype
        assert self.h == self.heads * 64, "kernels assume head_dim 64"

    def layer_range(self, mode):
        return (0, self.cfg["fusion_layer"]) if mode == "text" else (self.cfg["fusion_layer"],
                                                                      self.cfg["num_hidden_layers"])

    def _qkv(self, P, W, l):
        names = [l + f"attention.self.{n}." for n in ("query", "key", "value")]
        Wq = W.get_cat(l + "qkv.w", [P[n + "weight"] for n in names])
        bq = W.get_cat(l + "qkv.b", [P[n + "bias"] for n in names])
        return Wq, bq

    def _mask(self, shape, pdrop, seeds, dev):
        m = _empty(shape, self.dtype, dev)
        ops.dropout_mask(m, pdrop, next(seeds))
        return m

    def embed(self, P, ids, save, pdrop=0.0, seeds=None):
        """BertEmbeddings (xbert.py:186-213): gather-sum, LayerNorm, dropout. Returns x32, x16, ctx."""
        e = self.p + "bert.embeddings."
        B, L = ids.shape
        dev = ids.device
        h = self.h
        esum = _empty((B * L, h), torch.float32, dev)
        ops.bert_embed_gather(ids.contiguous(), P[e + "word_embeddings.weight"].detach(),
                              P[e + "position_embeddings.weight"].detach(),
                              P[e + "token_type_embeddings.weight"].detach(), esum, L, h)
        x32 = _empty((B * L, h), torch.float32, dev)
        x16 = _empty((B * L, h), self.dtype, dev)
        st = _empty((2, B * L), torch.float32, dev)
        mask = self._mask((B * L, h), pdrop, seeds, dev) if pdrop > 0 else None
        ops.layernorm_fwd(esum, P[e + "LayerNorm.weight"].detach(), P[e + "LayerNorm.bias"].detach(), self.eps,
                          out32=x32, out16=x16, mean=st[0], rstd=st[1], mul16=mask)
        return x32, x16, (dict(ids=ids, esum=esum, st=st, L=L, mask=mask) if save else None)

    def embed_backward(self, P, ctx, dx32, G, S):
        e = self.p + "bert.embeddings."
        h = self.h
        de = _empty(dx32.shape, torch.float32, dx32.device)
        ops.layernorm_bwd(dx32, ctx["esum"], ctx["st"][0], ctx["st"][1], P[e + "LayerNorm.weight"].detach(), de, 0,
                          dgamma=G[e + "LayerNorm.weight"], dbeta=G[e + "LayerNorm.bias"], param_scale=1.0 / S,
                          dy_mul16=ctx["mask"])
        ops.bert_embed_scatter(ctx["ids"].contiguous(), de, G[e + "word_embeddings.weight"],
                               G[e + "position_embeddings.weight"], G[e + "token_type_embeddings.weight"][0],
                               ctx["L"], h, 1.0 / S)

    def forward(self, P, W, x32, x16, add_mask, nseq, S_len, mode, save, pdrop=0.0, seeds=None, pattn=0.0):
        """x32/x16: [nseq*S_len, h]; add_mask fp32 [nseq, S_len]. Returns (y32, y16, ctx).
        pdrop > 0: train-mode hidden dropout after both dense output layers (xbert.py:358, 436)."""
        dev = x32.device
        h, heads, dt = self.h, self.heads, self.dtype
        M = nseq * S_len
        lo, hi = self.layer_range(mode)
        ctx = {"layers": [], "nseq": nseq, "S": S_len, "mode": mode, "mask": add_mask} if save else None
        scale = 1.0 / math.sqrt(64)
        for i in range(lo, hi):
            l = f"{self.p}bert.encoder.layer.{i}."
            g = lambda n: P[l + n].detach()
            w = lambda n: W.get(l + n, P[l + n])
            Wq, bq = self._qkv(P, W, l)
            qkv = _empty((M, 3 * h), dt, dev)
            ops.gemm16(x16, Wq, bias=bq, out16=qkv)
            cx = _empty((M, h), dt, dev)
            lse = _empty((nseq, heads, S_len), torch.float32, dev)
            aseed = next(seeds) if pattn > 0 else 0
            ops.seq_attn_fwd(qkv, add_mask, cx, None, lse, S_len, nseq, heads, 1, 1, S_len, scale, pattn, aseed)
            z1 = _empty((M, h), torch.float32, dev)
            mo = self._mask((M, h), pdrop, seeds, dev) if pdrop > 0 else None
            mf = self._mask((M, h), pdrop, seeds, dev) if pdrop > 0 else None
            ops.gemm16(cx, w("attention.output.dense.weight"), bias=g("attention.output.dense.bias"), resid=x32,
                       out32=z1, act=ops.ACT_GELU_GRAD if mo is not None else 0, aux=mo)   # MUL_AUX: dropout mask
            a32 = _empty((M, h), torch.float32, dev)
            a16 = _empty((M, h), dt, dev)
            st1 = _empty((2, M), torch.float32, dev)
            ops.layernorm_fwd(z1, g("attention.output.LayerNorm.weight"), g("attention.output.LayerNorm.bias"),
                              self.eps, out32=a32, out16=a16, mean=st1[0], rstd=st1[1])
            ff = self.cfg["intermediate_size"]
            hdn = _empty((M, ff), dt, dev)
            pre = _empty((M, ff), dt, dev) if save else None
            ops.gemm16(a16, w("intermediate.dense.weight"), bias=g("intermediate.dense.bias"), act=ACT_GELU, out16=hdn,
                       out16b=pre)
            z2 = _empty((M, h), torch.float32, dev)
            ops.gemm16(hdn, w("output.dense.weight"), bias=g("output.dense.bias"), resid=a32, out32=z2,
                       act=ops.ACT_GELU_GRAD if mf is not None else 0, aux=mf)
            y32 = _empty((M, h), torch.float32, dev)
            y16 = _empty((M, h), dt, dev)
            st2 = _empty((2, M), torch.float32, dev)
            ops.layernorm_fwd(z2, g("output.LayerNorm.weight"), g("output.LayerNorm.bias"), self.eps, out32=y32,
                              out16=y16, mean=st2[0], rstd=st2[1])
            if save:
                ctx["layers"].append(dict(i=i, x16=x16, qkv=qkv, cx=cx, lse=lse, z1=z1, st1=st1, a16=a16, hdn=hdn,
                                          pre=pre, z2=z2, st2=st2, mo=mo, mf=mf, pattn=pattn, aseed=aseed))
            x32, x16 = y32, y16
        return x32, x16, ctx

    def backward(self, P, W, ctx, dy32, G, S):
        """dy32: fp32 [M,h] gradient (scaled by S) of the encoder output. Returns gradient wrt the encoder input."""
        dev = dy32.device
        h, heads, dt = self.h, self.heads, self.dtype
        nseq, S_len = ctx["nseq"], ctx["S"]
        M = nseq * S_len
        inv = 1.0 / S
        scale = 1.0 / math.sqrt(64)
        ff = self.cfg["intermediate_size"]

        def wgrad(dy16, x16, gw, gb=None):
            ops.gemm16(dy16, x16, a_layout=MNMAJOR, b_layout=MNMAJOR, out32=gw, split_k=-1, alpha=inv)
            if gb is not None:
                ops.colsum(dy16, gb, inv)

        for c in reversed(ctx["layers"]):
            l = f"{self.p}bert.encoder.layer.{c['i']}."
            g = lambda n: P[l + n].detach()
            w = lambda n: W.get(l + n, P[l + n])
            dz2 = _empty((M, h), torch.float32, dev)
            dz2_16 = _empty((M, h), dt, dev)
            ops.layernorm_bwd(dy32, c["z2"], c["st2"][0], c["st2"][1], g("output.LayerNorm.weight"), dz2, 0,
                              dx16=dz2_16, dgamma=G[l + "output.LayerNorm.weight"],
                              dbeta=G[l + "output.LayerNorm.bias"], param_scale=inv, colsum=G[l + "output.dense.bias"],
                              dx16_mul16=c["mf"])
            du = _empty((M, ff), dt, dev)
            ops.gemm16(dz2_16, w("output.dense.weight"), b_layout=MNMAJOR, act=ACT_GELU_GRAD, aux=c["pre"], out16=du)
            wgrad(dz2_16, c["hdn"], G[l + "output.dense.weight"])
            da32 = _empty((M, h), torch.float32, dev)
            ops.gemm16(du, w("intermediate.dense.weight"), b_layout=MNMAJOR, resid=dz2, out32=da32)
            wgrad(du, c["a16"], G[l + "intermediate.dense.weight"], G[l + "intermediate.dense.bias"])
            dz1 = dz2  # reuse
            dz1_16 = dz2_16
            ops.layernorm_bwd(da32, c["z1"], c["st1"][0], c["st1"][1], g("attention.output.LayerNorm.weight"), dz1, 0,
                              dx16=dz1_16, dgamma=G[l + "attention.output.LayerNorm.weight"],
                              dbeta=G[l + "attention.output.LayerNorm.bias"], param_scale=inv,
                              colsum=G[l + "attention.output.dense.bias"], dx16_mul16=c["mo"])
            dcx = _empty((M, h), dt, dev)
            ops.gemm16(dz1_16, w("attention.output.dense.weight"), b_layout=MNMAJOR, out16=dcx)
            wgrad(dz1_16, c["cx"], G[l + "attention.output.dense.weight"])
            dqkv = _empty((M, 3 * h), dt, dev)
            ops.seq_attn_bwd(c["qkv"], ctx["mask"], c["lse"], c["cx"], None, dcx, dqkv, None, S_len, nseq, heads, 1, 1,
                             S_len, scale, drop_p=c["pattn"], drop_seed=c["aseed"])
            Wq, _ = self._qkv(P, W, l)
            dx32 = da32  # reuse
            ops.gemm16(dqkv, Wq, b_layout=MNMAJOR, resid=dz1, out32=dx32)
            wgrad(dqkv, c["x16"], G.group(l + "qkv.w"), G.group(l + "qkv.b"))
            dy32 = dx32
        return dy32


# =====================================================================================================================
class LocalComm:
    """World-size-1 stand-in for the VTC feature exchange (hvd.allgather, alpro_models.py:110-111)."""
    rank = 0
    world = 1

    def all_gather(self, x):
        return x

    def reduce_scatter_sum(self, g):
        return g


def multinomial_sampler(weights):
    """Row-wise multinomial draw through torch (torch.multinomial(w[b], 1) per row in the reference,
    alpro_models.py:301-316). Kept as an alternative `engine.sampler`; the default is the fused Philox kernel."""
    return torch.multinomial(weights, 1).squeeze(1)


def argmax_sampler(weights):
    """Deterministic rule used on both sides of every parity test."""
    return torch.argmax(weights, dim=1)


class AlproEngine:
    """Forward + backward of AlproForVideoTextRetrieval / AlproForPretrain / Prompter features on the CUDA kernels."""

    def __init__(self, kind, bert_cfg, vis, dtype=torch.float16, loss_scale=4096.0, num_entities=None):
        assert kind in ("retrieval", "pretrain", "prompter")
        self.kind = kind
        self.cfg = bert_cfg
        self.vis = vis
        self.dtype = dtype
        self.S = float(loss_scale) if dtype == torch.float16 else 1.0
        self.visual = VisualEncoder("visual_encoder.model.", vis, dtype)
        self.bert = BertEncoder("text_encoder.", bert_cfg, dtype)
        self.W = OperandCache(dtype)
        self.num_entities = num_entities
        if kind == "pretrain":
            self.t_visual = VisualEncoder("prompter.visual_encoder.model.", vis, dtype)
        self.sampler = None              # None: fused Philox draw; or a callable on the [B,B] weights (tests: argmax)
        self.neg_seed = 0x5851F42D4C957F2D
        self._draw = 0
        self.comm = LocalComm()
        self.last_grads = None
        self.base_seed = 0x5DEECE66
        self._step = 0
        self.grad_ready_hook = None

    # ------------------------------------------------------------------------------------------------ features
    def _proj_norm(self, P, x, ldx, wname, rows):
        """F.normalize(Linear(x[:,0,:])) (alpro_models.py:103,205): x rows are `ldx` floats apart."""
        Wt, bt = P[wname + ".weight"].detach(), P[wname + ".bias"].detach()
        K = Wt.shape[1]
        proj = _empty((rows, 256), torch.float32, x.device)
        ops.small_linear_fwd(x, ldx, Wt, bt, proj, rows, 256, K)
        feat = _empty((rows, 256), torch.float32, x.device)
        nrm = _empty((rows,), torch.float32, x.device)
        ops.l2norm_fwd(proj, feat, nrm)
        return feat, nrm

    def _text_mask_add(self, mask):
        return ((1.0 - mask.to(torch.float32)) * -10000.0).contiguous()

    def clamp_temp(self, P):
        ops.clamp_scalar(P["temp"].detach(), 0.001, 0.5)                       # temp.clamp_ :80-81 / :598-599 / :734-735

    def text_features(self, P, ids, mask):
        """No-grad text pass: (text_embeds [n,L,h], F.normalize(text_proj(cls)) [n,256]); alpro_models.py:196-207."""
        ids, mask = ids.contiguous(), mask.contiguous()
        n, L = ids.shape
        h = self.cfg["hidden_size"]
        x32, x16, _ = self.bert.embed(P, ids, False)
        te, _, _ = self.bert.forward(P, self.W, x32, x16, self._text_mask_add(mask), n, L, "text", False)
        tf, _ = self._proj_norm(P, te, L * h, "text_proj", n)
        return te.view(n, L, h), tf

    def visual_feat(self, P, frames):
        """No-grad visual pass: (video_embeds [B,1+N,d], F.normalize(vision_proj(cls)) [B,256]); :509-523."""
        ve, _ = self.visual.forward(P, self.W, frames, False)
        vf, _ = self._proj_norm(P, ve, ve.shape[1] * self.vis["d"], "vision_proj", ve.shape[0])
        return ve, vf

    def pseudo_labels(self, P, frames, type_="video"):
        """Prompter.get_pseudo_labels (alpro_models.py:531-551) for a Prompter-kind engine: softmax of the similarity
        to the prompt features / temp; ignore iff the argmax INDEX < 0.2 (sic, :527)."""
        B = frames.shape[0]
        _, vf = self.visual_feat(P, frames)
        prompt = P["video_prompt_feat"] if type_ == "video" else P["image_prompt_feat"]
        E = prompt.shape[0]
        sim = _empty((B, E), torch.float32, frames.device)
        ops.small_linear_fwd(vf, 256, prompt.detach().contiguous(), None, sim, B, E, 256, 1.0, P["temp"].detach(), 2)
        soft = _empty((B, E), torch.float32, frames.device)
        ignore = _empty((B,), torch.uint8, frames.device)
        ops.pseudo_labels(sim, soft, ignore)
        return soft, ignore.bool()

    # ------------------------------------------------------------------------------------------------ forward
    def _seed_stream(self):
        """Python-side stream of 32-bit seeds for the dropout sites of one forward pass (no device sync)."""
        self._step += 1
        base = (self.base_seed * 0x9E3779B1 + self._step * 0x85EBCA77) & 0xFFFFFFFF
        site = 0
        while True:
            site += 1
            yield (base ^ (site * 0xC2B2AE3D)) & 0xFFFFFFFF

    def forward(self, P, batch, need_grad=True, training=False):
        """Returns (outputs, ctx). outputs mirrors the reference dict (alpro_models.py:172-183, 793-798).
        training=True applies the reference's train-mode regularisers: BERT hidden dropout (xbert.py:178,358,436) and
        attention-probability dropout (xbert.py:331), TimeSformer stochastic depth (vit.py:157,181,212)."""
        kind = self.kind
        dev = batch["visual_inputs"].device
        cfg, h, d = self.cfg, self.cfg["hidden_size"], self.vis["d"]
        save = need_grad
        comm = self.comm
        if need_grad:
            self.W.refresh()         # a training step never trusts operand copies made before it (see OperandCache)
        ops.clamp_scalar(P["temp"].detach(), 0.001, 0.5)                       # temp.clamp_ :80-81 / :734-735
        frames = batch["visual_inputs"]
        B = frames.shape[0]
        pdrop = float(cfg.get("hidden_dropout_prob", 0.0)) if training else 0.0
        dpr = float(self.vis.get("drop_path_rate", 0.0)) if training else 0.0
        pattn = float(cfg.get("attention_probs_dropout_prob", 0.0)) if training else 0.0
        seeds = self._seed_stream() if (pdrop > 0 or pattn > 0) else None
        ve, vctx = self.visual.forward(P, self.W, frames, save, drop_path_rate=dpr)     # [B, Nv, d]
        Nv = ve.shape[1]
        ids, mask = batch["text_input_ids"], batch["text_input_mask"]
        L = ids.shape[1]
        use_mlm = kind == "pretrain" and "mlm_labels" in batch
        use_mpm = kind == "pretrain" and "mpm_mask" in batch
        if use_mlm:
            ids_all = torch.cat([ids, batch["mlm_text_input_ids"]], dim=0)
            mask_all = torch.cat([mask, mask], dim=0)
        else:
            ids_all, mask_all = ids, mask
        nt = ids_all.shape[0]
        mask_all = mask_all.contiguous()
        x32, x16, ectx = self.bert.embed(P, ids_all, save, pdrop, seeds)
        te, _, tctx = self.bert.forward(P, self.W, x32, x16, self._text_mask_add(mask_all), nt, L, "text", save, pdrop,
                                        seeds, pattn)
        te = te.view(nt, L, h)

        # ---- VTC (alpro_models.py:103-128, 750-779)
        vf, vnorm = self._proj_norm(P, ve, Nv * d, "vision_proj", B)
        tf, tnorm = self._proj_norm(P, te, L * h, "text_proj", B)
        # ONE exchange for both feature sets (the reference issues two hvd.allgather calls, alpro_models.py:110-111):
        # [B, 512] = video | text rows, gathered in rank order; gv / gt are strided views into the gathered buffer
        if comm.world > 1:
            gvt = comm.all_gather(torch.cat([vf, tf], dim=1))
            gv, gt, ldg = gvt[:, :256], gvt[:, 256:], 512
        else:
            gv, gt, ldg = vf, tf, 256
        Gn = gv.shape[0]
        temp = P["temp"].detach()
        sim_v2t = _empty((B, Gn), torch.float32, dev)
        sim_t2v = _empty((B, Gn), torch.float32, dev)
        ops.small_linear_fwd(vf, 256, gt, None, sim_v2t, B, Gn, 256, 1.0, temp, 2, ldw=ldg)
        ops.small_linear_fwd(tf, 256, gv, None, sim_t2v, B, Gn, 256, 1.0, temp, 2, ldw=ldg)
        vtc_labels = torch.arange(B, device=dev, dtype=torch.int64) + B * comm.rank   # local_rank block :119-123
        ce_v = ops.softmax_ce_fwd(sim_v2t, Gn, hard=vtc_labels, denom_mode=1)
        ce_t = ops.softmax_ce_fwd(sim_t2v, Gn, hard=vtc_labels, denom_mode=1)
        itc_loss = (ce_v.loss + ce_t.loss) * 0.5

        if kind == "prompter":
            # Prompter.forward (alpro_models.py:553-595): the contrastive objective alone
            out = dict(itc_loss=itc_loss, itc_labels=vtc_labels, i2t_scores=None, t2i_scores=None)
            lsm_v = sim_v2t - ce_v.row_lse.view(B, 1)                          # F.log_softmax(sim, dim=1) :581-582
            lsm_t = sim_t2v - ce_t.row_lse.view(B, 1)
            out.update(i2t_scores=lsm_v, t2i_scores=lsm_t, _video_embeds=ve, _text_embeds=te[:B])
            ctx = None
            if save:
                ctx = dict(B=B, L=L, Nv=Nv, nt=nt, ve=ve, te=te, vctx=vctx, ectx=ectx, tctx=tctx, vf=vf, vnorm=vnorm,
                           tf=tf, tnorm=tnorm, gv=gv, gt=gt, sim_v2t=sim_v2t, sim_t2v=sim_t2v, ce_v=ce_v, ce_t=ce_t,
                           vtc_labels=vtc_labels, use_mlm=False, use_mpm=False, vtc_only=True)
            self.last_ctx = ctx
            return out, ctx

        # ---- hard negatives (alpro_models.py:288-316)
        if B <= 1:
            neg_video = neg_text = torch.zeros(1, dtype=torch.int64, device=dev)
        elif self.sampler is None:
            # production path: weights + the per-row multinomial draw in ONE kernel per direction (Philox4x32-10 keyed by
            # (neg_seed, step); the reference draws row by row with 2*B .item() synchronisations)
            self._draw += 2
            neg_video = _empty((B,), torch.int64, dev)
            neg_text = _empty((B,), torch.int64, dev)
            ops.neg_sample(sim_t2v, B * comm.rank, B, self.neg_seed, self._draw, neg_video)      # a negative video per text
            ops.neg_sample(sim_v2t, B * comm.rank, B, self.neg_seed, self._draw + 1, neg_text)   # a negative text per video
        else:
            w_t2v = _empty((B, B), torch.float32, dev)
            w_v2t = _empty((B, B), torch.float32, dev)
            ops.neg_weights(sim_t2v, B * comm.rank, B, w_t2v)
            ops.neg_weights(sim_v2t, B * comm.rank, B, w_v2t)
            neg_video = self.sampler(w_t2v)    # a negative video for each text
            neg_text = self.sampler(w_v2t)     # a negative text for each video

        # ---- one batched fusion pass: positives | (text_i, video_neg_i) | (text_neg_i, video_i) | MLM pairs
        ar = torch.arange(B, device=dev, dtype=torch.int64)
        ti = [ar, ar, neg_text]
        vi = [ar, neg_video, ar]
        if use_mlm:
            ti.append(ar + B)
            vi.append(ar)
        ti = torch.cat(ti).to(torch.int32).contiguous()
        vi = torch.cat(vi).to(torch.int32).contiguous()
        S_all = ti.numel()
        R = L + Nv
        f32 = _empty((S_all * R, h), torch.float32, dev)
        f16 = _empty((S_all * R, h), self.dtype, dev)
        fmask = _empty((S_all, R), torch.float32, dev)
        ops.fusion_gather_fwd(te, ve, mask_all, ti, vi, f32, f16, fmask, S_all, L, Nv, h)
        fo, _, fctx = self.bert.forward(P, self.W, f32, f16, fmask, S_all, R, "fusion", save, pdrop, seeds, pattn)

        # ---- VTM head (alpro_models.py:334-339)
        itm_scores = _empty((3 * B, 2), torch.float32, dev)
        ops.small_linear_fwd(fo, R * h, P["itm_head.weight"].detach(), P["itm_head.bias"].detach(), itm_scores,
                             3 * B, 2, h)
        itm_labels = torch.cat([torch.ones(B, dtype=torch.int64, device=dev),
                                torch.zeros(2 * B, dtype=torch.int64, device=dev)])
        ce_itm = ops.softmax_ce_fwd(itm_scores, 2, hard=itm_labels, denom_mode=1)

        out = dict(itm_scores=itm_scores, itm_loss=ce_itm.loss, itm_labels=itm_labels, itc_loss=itc_loss)
        ctx = None
        if save:
            ctx = dict(B=B, L=L, Nv=Nv, R=R, S_all=S_all, nt=nt, ve=ve, te=te, vctx=vctx, ectx=ectx, tctx=tctx,
                       fctx=fctx, fo=fo, ti=ti, vi=vi, vf=vf, vnorm=vnorm, tf=tf, tnorm=tnorm, gv=gv, gt=gt,
                       sim_v2t=sim_v2t, sim_t2v=sim_t2v, ce_v=ce_v, ce_t=ce_t, vtc_labels=vtc_labels,
                       itm_scores=itm_scores, itm_labels=itm_labels, ce_itm=ce_itm, use_mlm=use_mlm, use_mpm=use_mpm)
        self.last_ctx = ctx   # introspection hook for the train-mode parity tests (regulariser masks live in ctx)
        out["_neg_video"], out["_neg_text"] = neg_video, neg_text
        out["_video_embeds"], out["_text_embeds"] = ve, te[:B]

        if kind == "pretrain":
            out.update(mlm_scores=None, mlm_loss=None, mlm_labels=None, mpm_loss=None, mpm_logits=None, mpm_labels=None)
        if use_mlm:
            self._mlm_forward(P, batch, fo, B, L, R, h, out, ctx)
        if use_mpm:
            self._mpm_forward(P, batch, fo, B, L, Nv, R, h, out, ctx)
        return out, ctx

    def _mlm_forward(self, P, batch, fo, B, L, R, h, out, ctx):
        """compute_mlm head part (alpro_models.py:366-371) on the MLM sequences [3B, 4B) of the fusion batch;
        BertLMPredictionHead xbert.py:648-682."""
        dev = fo.device
        c = "text_encoder.cls.predictions."
        V = self.cfg["vocab_size"]
        M = B * L
        tin16 = _empty((M, h), self.dtype, dev)
        ops.take_rows_fwd(fo, R, 3 * B, B, L, h, out16=tin16)
        g32 = _empty((M, h), torch.float32, dev)
        pre16 = _empty((M, h), self.dtype, dev)
        ops.gemm16(tin16, self.W.get(c + "transform.dense.weight", P[c + "transform.dense.weight"]),
                   bias=P[c + "transform.dense.bias"].detach(), act=ACT_GELU, out32=g32, out16b=pre16)
        t16 = _empty((M, h), self.dtype, dev)
        st = _empty((2, M), torch.float32, dev)
        ops.layernorm_fwd(g32, P[c + "transform.LayerNorm.weight"].detach(), P[c + "transform.LayerNorm.bias"].detach(),
                          self.cfg["layer_norm_eps"], out16=t16, mean=st[0], rstd=st[1])
        wname = "text_encoder.bert.embeddings.word_embeddings.weight"
        Wdec = self.W.get(wname, P[wname])
        logits = _empty((M, V), torch.float32, dev)
        ops.gemm16(t16, Wdec, bias=P[c + "bias"].detach(), out32=logits)
        labels = batch["mlm_labels"].contiguous().view(-1)
        ce = ops.softmax_ce_fwd(logits, V, hard=labels, denom_mode=0)
        out.update(mlm_scores=logits.view(B, L, V), mlm_loss=ce.loss, mlm_labels=batch["mlm_labels"])
        if ctx is not None:
            ctx.update(mlm=dict(tin16=tin16, g32=g32, pre16=pre16, t16=t16, st=st, logits=logits, labels=labels, ce=ce))

    def _mpm_forward(self, P, batch, fo, B, L, Nv, R, h, out, ctx):
        """Prompter.get_pseudo_labels (teacher, no grad; alpro_models.py:531-551) + compute_mpm_with_encoder_out
        (:209-232) on the positive fusion outputs."""
        dev = fo.device
        d = self.vis["d"]
        E = self.num_entities
        Pp = {k[len("prompter."):]: v for k, v in P.items() if k.startswith("prompter.")}
        tve, _ = VisualEncoder("visual_encoder.model.", self.vis, self.dtype).forward(
            Pp, _PrefixedCache(self.W, "prompter."), batch["crop_visual_inputs"], False)
        tproj = _empty((B, 256), torch.float32, dev)
        ops.small_linear_fwd(tve, tve.shape[1] * d, Pp["vision_proj.weight"].detach(), Pp["vision_proj.bias"].detach(),
                             tproj, B, 256, d)
        tfeat = _empty((B, 256), torch.float32, dev)
        tn = _empty((B,), torch.float32, dev)
        ops.l2norm_fwd(tproj, tfeat, tn)
        prompt = Pp["video_prompt_feat"] if batch.get("type", "video") == "video" else Pp["image_prompt_feat"]
        sim = _empty((B, E), torch.float32, dev)
        ops.small_linear_fwd(tfeat, 256, prompt.detach().contiguous(), None, sim, B, E, 256, 1.0, Pp["temp"].detach(), 2)
        soft = _empty((B, E), torch.float32, dev)
        ignore = _empty((B,), torch.uint8, dev)
        ops.pseudo_labels(sim, soft, ignore)
        # student
        pm = batch["mpm_mask"].to(torch.float32).contiguous().view(B, -1)
        Np = pm.shape[1]
        pooled = _empty((B, h), torch.float32, dev)
        ops.masked_mean_fwd(fo, R * h, L + 1, pm, B, Np, h, pooled)
        h1 = _empty((B, 2 * h), torch.float32, dev)
        ops.small_linear_fwd(pooled, h, P["mpm_head.0.weight"].detach(), P["mpm_head.0.bias"].detach(), h1, B, 2 * h, h,
                             relu=True)
        logits = _empty((B, E), torch.float32, dev)
        ops.small_linear_fwd(h1, 2 * h, P["mpm_head.2.weight"].detach(), P["mpm_head.2.bias"].detach(), logits, B, E,
                             2 * h)
        ce = ops.softmax_ce_fwd(logits, E, soft=soft, row_ignore=ignore, denom_mode=0)
        out.update(mpm_loss=ce.loss, mpm_logits=logits, mpm_labels=soft, _mpm_ignore=ignore.bool())
        if ctx is not None:
            ctx.update(mpm=dict(pm=pm, Np=Np, pooled=pooled, h1=h1, logits=logits, soft=soft, ce=ce))

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, P, ctx, named_params, g):
        """g: dict loss name -> device scalar (upstream gradient) or None. Returns a GradStore over named_params."""
        S = self.S
        inv = 1.0 / S
        dev = ctx["ve"].device
        cfg, h, d = self.cfg, self.cfg["hidden_size"], self.vis["d"]
        vtc_only = bool(ctx.get("vtc_only"))
        B, L, Nv, nt = ctx["B"], ctx["L"], ctx["Nv"], ctx["nt"]
        R, S_all = (0, 0) if vtc_only else (ctx["R"], ctx["S_all"])
        alloc = None
        if getattr(self, "grad_alloc", None) is not None:
            live = getattr(self, "grad_accum", None)            # a store still aliased by p.grad must not be recycled
            alloc = lambda n, d: self.grad_alloc(n, d, avoid=live.flat if live is not None else None)
        G = GradStore(named_params, dev, bert_grad_groups("text_encoder.", cfg), alloc=alloc)
        comm = self.comm
        dfo = None if vtc_only else torch.zeros(S_all * R, h, dtype=torch.float32, device=dev)   # d fusion output (scaled)

        # ---- heads on top of the fusion encoder
        if ctx["use_mpm"] and g.get("mpm_loss") is not None:
            m = ctx["mpm"]
            E = self.num_entities
            dlog = _empty((B, E), torch.float32, dev)
            ops.softmax_ce_bwd(m["logits"], E, m["ce"], g["mpm_loss"], S, soft=m["soft"], out32=dlog)
            dh1 = _empty((B, 2 * h), torch.float32, dev)
            ops.small_linear_bwd(dlog, E, None, m["h1"], 2 * h, P["mpm_head.2.weight"].detach(), dh1, 2 * h, 0,
                                 G["mpm_head.2.weight"], G["mpm_head.2.bias"], 0, B, E, 2 * h, dw_scale=inv)
            dpool = _empty((B, h), torch.float32, dev)
            ops.small_linear_bwd(dh1, 2 * h, m["h1"], m["pooled"], h, P["mpm_head.0.weight"].detach(), dpool, h, 0,
                                 G["mpm_head.0.weight"], G["mpm_head.0.bias"], 0, B, 2 * h, h, dw_scale=inv)
            ops.masked_mean_bwd(dpool, m["pm"], B, m["Np"], h, dfo, R * h, L + 1)
        if ctx["use_mlm"] and g.get("mlm_loss") is not None:
            self._mlm_backward(P, ctx, G, g["mlm_loss"], dfo)
        if not vtc_only and g.get("itm_loss") is not None:
            dlog = _empty((3 * B, 2), torch.float32, dev)
            ops.softmax_ce_bwd(ctx["itm_scores"], 2, ctx["ce_itm"], g["itm_loss"], S, hard=ctx["itm_labels"], out32=dlog)
            ops.small_linear_bwd(dlog, 2, None, ctx["fo"], R * h, P["itm_head.weight"].detach(), dfo, R * h, 1,
                                 G["itm_head.weight"], G["itm_head.bias"], 0, 3 * B, 2, h, dw_scale=inv)

        # ---- fusion encoder and the gather that built its input
        dte = torch.zeros(nt, L, h, dtype=torch.float32, device=dev)
        dve = torch.zeros(B, Nv, d, dtype=torch.float32, device=dev)
        if not vtc_only:
            dfi = self.bert.backward(P, self.W, ctx["fctx"], dfo, G, S)
            ops.fusion_gather_bwd(dfi, ctx["ti"], ctx["vi"], dte, dve, S_all, L, Nv, h)

        # ---- VTC
        if g.get("itc_loss") is not None:
            Gn = ctx["gv"].shape[0]
            temp = P["temp"].detach()
            ds_v = _empty((B, Gn), torch.float32, dev)
            ds_t = _empty((B, Gn), torch.float32, dev)
            ops.softmax_ce_bwd(ctx["sim_v2t"], Gn, ctx["ce_v"], g["itc_loss"], 0.5 * S, hard=ctx["vtc_labels"], out32=ds_v)
            ops.softmax_ce_bwd(ctx["sim_t2v"], Gn, ctx["ce_t"], g["itc_loss"], 0.5 * S, hard=ctx["vtc_labels"], out32=ds_t)
            dvf = _empty((B, 256), torch.float32, dev)
            dtf = _empty((B, 256), torch.float32, dev)
            ldg = ctx["gv"].stride(0)
            dgvt = _empty((Gn, 512), torch.float32, dev)        # d gathered (video | text), same layout as the exchange
            dgv, dgt = dgvt[:, :256], dgvt[:, 256:]
            # sim_v2t = vf gt^T / temp : d vf = ds_v gt / temp ; d gt = ds_v^T vf / temp
            ops.small_linear_bwd(ds_v, Gn, None, ctx["vf"], 256, ctx["gt"], dvf, 256, 0, dgt, None, 0, B, Gn, 256,
                                 alpha=1.0, alpha_dev=temp, alpha_mode=2, ldw=ldg, lddw=512)
            ops.small_linear_bwd(ds_t, Gn, None, ctx["tf"], 256, ctx["gv"], dtf, 256, 0, dgv, None, 0, B, Gn, 256,
                                 alpha=1.0, alpha_dev=temp, alpha_mode=2, ldw=ldg, lddw=512)
            # backward of the all-gather: sum over ranks, keep the local slice (Horovod allgather grad semantics);
            # one reduce-scatter for both feature sets
            if comm.world > 1:
                loc = comm.reduce_scatter_sum(dgvt)
                dvf.add_(loc[:, :256])
                dtf.add_(loc[:, 256:])
            else:
                dvf.add_(dgv)
                dtf.add_(dgt)
            ops.temp_grad(ds_v, ctx["sim_v2t"], ds_t, ctx["sim_t2v"], temp, G["temp"].view(1), inv)
            for feat, nrm, dfeat, wname, x, ldx, dx in ((ctx["vf"], ctx["vnorm"], dvf, "vision_proj", ctx["ve"], Nv * d, dve),
                                                        (ctx["tf"], ctx["tnorm"], dtf, "text_proj", ctx["te"], L * h, dte)):
                dproj = _empty((B, 256), torch.float32, dev)
                ops.l2norm_bwd(dfeat, feat, nrm, dproj)
                K = P[wname + ".weight"].shape[1]
                ops.small_linear_bwd(dproj, 256, None, x, ldx, P[wname + ".weight"].detach(), dx, ldx, 1,
                                     G[wname + ".weight"], G[wname + ".bias"], 0, B, 256, K, dw_scale=inv)

        # ---- encoders
        dx_text = self.bert.backward(P, self.W, ctx["tctx"], dte.view(nt * L, h), G, S)
        self.bert.embed_backward(P, ctx["ectx"], dx_text, G, S)
        self.last_grads = G
        hook = self.grad_ready_hook                    # data-parallel overlap: called with the flat prefix that is final
        if hook is not None:
            hook(G, G.region_end(("nonvis",)))
        self.visual.backward(P, self.W, ctx["vctx"], dve, G, S, hook)
        if hook is not None:
            hook(G, G.flat.numel())
        self.W.invalidate()          # an optimizer step (possibly through p.data) follows: next forward re-casts
        return G

    def _mlm_backward(self, P, ctx, G, gptr, dfo):
        S, inv = self.S, 1.0 / self.S
        m = ctx["mlm"]
        B, L, R = ctx["B"], ctx["L"], ctx["R"]
        h = self.cfg["hidden_size"]
        V = self.cfg["vocab_size"]
        dev = dfo.device
        c = "text_encoder.cls.predictions."
        M = B * L
        Vp = (V + 7) // 8 * 8
        dlog = _empty((M, Vp), self.dtype, dev)
        ops.softmax_ce_bwd(m["logits"], V, m["ce"], gptr, S, hard=m["labels"], out16=dlog, C_out=Vp)
        dlv = dlog[:, :V]
        wname = "text_encoder.bert.embeddings.word_embeddings.weight"
        Wdec = self.W.get(wname, P[wname])
        dt16 = _empty((M, h), self.dtype, dev)
        ops.gemm16(dlv, Wdec, b_layout=MNMAJOR, out16=dt16)
        ops.gemm16(dlv, m["t16"], a_layout=MNMAJOR, b_layout=MNMAJOR, out32=G[wname], split_k=-1, alpha=inv)
        ops.colsum(dlv, G[c + "bias"], inv)
        dg32 = _empty((M, h), torch.float32, dev)
        ops.layernorm_bwd(dt16, m["g32"], m["st"][0], m["st"][1], P[c + "transform.LayerNorm.weight"].detach(), dg32, 0,
                          dgamma=G[c + "transform.LayerNorm.weight"], dbeta=G[c + "transform.LayerNorm.bias"],
                          param_scale=inv)
        du16 = _empty((M, h), self.dtype, dev)
        ops.gelu_grad_mul(dg32, m["pre16"], du16)
        dtin = _empty((M, h), torch.float32, dev)
        ops.gemm16(du16, self.W.get(c + "transform.dense.weight", P[c + "transform.dense.weight"]), b_layout=MNMAJOR,
                   out32=dtin)
        ops.gemm16(du16, m["tin16"], a_layout=MNMAJOR, b_layout=MNMAJOR, out32=G[c + "transform.dense.weight"],
                   split_k=-1, alpha=inv)
        ops.colsum(du16, G[c + "transform.dense.bias"], inv)
        ops.take_rows_bwd(dtin, R, 3 * B, B, L, h, dfo)

    # ------------------------------------------------------------------------------------------------ inference
    def inference(self, P, batch):
        """AlproForVideoTextRetrieval.forward_inference (alpro_models.py:874-914): 1 video x n texts. The video is
        encoded once and indexed n times by the fusion gather instead of `video_embeds.repeat(n, 1, 1)`."""
        dev = batch["visual_inputs"].device
        h, d = self.cfg["hidden_size"], self.vis["d"]
        ve, _ = self.visual.forward(P, self.W, batch["visual_inputs"], False)
        nvid, Nv = ve.shape[0], ve.shape[1]
        ids, mask = batch["text_input_ids"], batch["text_input_mask"].contiguous()
        n, L = ids.shape
        x32, x16, _ = self.bert.embed(P, ids, False)
        te, _, _ = self.bert.forward(P, self.W, x32, x16, self._text_mask_add(mask), n, L, "text", False)
        vf, _ = self._proj_norm(P, ve, Nv * d, "vision_proj", nvid)
        tf, _ = self._proj_norm(P, te, L * h, "text_proj", n)
        itc = _empty((nvid, n), torch.float32, dev)
        ops.small_linear_fwd(vf, 256, tf, None, itc, nvid, n, 256, 1.0, P["temp"].detach(), 2)
        ti = torch.arange(n, device=dev, dtype=torch.int32)
        vi = torch.zeros(n, device=dev, dtype=torch.int32) if nvid == 1 else torch.arange(n, device=dev, dtype=torch.int32)
        R = L + Nv
        f32 = _empty((n * R, h), torch.float32, dev)
        f16 = _empty((n * R, h), self.dtype, dev)
        fmask = _empty((n, R), torch.float32, dev)
        ops.fusion_gather_fwd(te, ve, mask, ti, vi, f32, f16, fmask, n, L, Nv, h)
        fo, _, _ = self.bert.forward(P, self.W, f32, f16, fmask, n, R, "fusion", False)
        logits = _empty((n, 2), torch.float32, dev)
        ops.small_linear_fwd(fo, R * h, P["itm_head.weight"].detach(), P["itm_head.bias"].detach(), logits, n, 2, h)
        return dict(logits=logits, itc_scores=itc)

    def visual_features(self, P, frames):
        return self.visual.forward(P, self.W, frames, False)[0]


class _PrefixedCache:
    """OperandCache view that namespaces entries (the frozen teacher copy lives under 'prompter.')."""

    def __init__(self, cache, prefix):
        self.cache, self.prefix = cache, prefix

    def get(self, name, p, shape2d=None):
        return self.cache.get(self.prefix + name, p, shape2d)

    def get_cat(self, name, ps):
        return self.cache.get_cat(self.prefix + name, ps)

    def get_fused_linear(self, name, w2, w1, b1):
        return self.cache.get_fused_linear(self.prefix + name, w2, w1, b1)
