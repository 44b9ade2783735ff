"""CPU: the C-ABI library loads and exports every symbol declared in include/alpro_b200.h (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from alpro_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built):
    from alpro_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "alpro_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    declared = set(re.findall(r"\b(alpro_\w+)\s*\(", hdr))
    assert len(declared) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (alpro_\w+)", out))
    assert declared <= exported, declared - exported
    assert set(_lib.PROTOS) == declared
    lib = ctypes.CDLL(built)
    for name in declared:
        getattr(lib, name)
    assert _lib.lib.alpro_version() >= 1


def test_header_is_plain_c(built):
    r = subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", os.path.join(ROOT, "include", "alpro_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_epilogue_struct_layout_matches_header(built):
    """ctypes mirror of AlproGemmEpilogue must have the C layout (checked by compiling a sizeof/offsetof probe)."""
    from alpro_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "alpro_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(AlproGemmEpilogue), offsetof(AlproGemmEpilogue, ld32),
         offsetof(AlproGemmEpilogue, out16_fmt), offsetof(AlproGemmEpilogue, act),
         offsetof(AlproGemmEpilogue, split_k), offsetof(AlproGemmEpilogue, alpha),
         offsetof(AlproGemmEpilogue, row_scale_bias));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(x) for x in subprocess.check_output([exe], text=True).split()]
    E = _lib.GemmEpilogue
    assert vals == [ctypes.sizeof(E), E.ld32.offset, E.out16_fmt.offset, E.act.offset, E.split_k.offset, E.alpha.offset,
                    E.row_scale_bias.offset]


def test_modules_keep_reference_state_dict_schema():
    """state_dict names/shapes of the nn.Module mirrors == alpro_b200.synth.model_spec (itself checked against the
    reference in tests/test_oracle_golden.py::test_state_dict_schema_matches_reference)."""
    from alpro_b200 import modeling, synth
    from oracle import configs
    for kind, cls in (("retrieval", modeling.AlproForVideoTextRetrieval), ("pretrain", modeling.AlproForPretrain)):
        cfg = configs.tiny(kind)
        v = dict(cfg["video"])
        v.update(embed_dim=192, depth=2, num_heads=3)
        b = dict(cfg["bert"])
        b["num_entities"] = cfg["num_entities"]
        m = cls(b, v)
        spec = synth.model_spec(kind, cfg["bert"], cfg["vis"], cfg["num_entities"])
        sd = m.state_dict()
        assert {k: tuple(t.shape) for k, t in sd.items()} == {k: tuple(s) for k, s in spec.items()}
        te = m.text_encoder
        assert te.cls.predictions.decoder.weight is te.bert.embeddings.word_embeddings.weight
        assert te.cls.predictions.decoder.bias is te.cls.predictions.bias
        m.load_state_dict(synth.synth_state_dict(spec, 3), strict=True)
        names = [n for n, _ in m.named_parameters()]
        assert "text_encoder.cls.predictions.decoder.weight" not in names   # tied alias is not a separate parameter
        if kind == "pretrain":
            assert all(not p.requires_grad for n, p in m.named_parameters() if n.startswith("prompter."))


def test_product_path_does_not_import_oracle():
    import ast
    pkg = os.path.join(ROOT, "alpro_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, f)).read())
            for node in ast.walk(tree):
                if isinstance(node, (ast.Import, ast.ImportFrom)):
                    mod = getattr(node, "module", None) or ""
                    names = [a.name for a in node.names]
                    assert not mod.startswith("oracle") and not any(n.startswith("oracle") for n in names), f
