"""The public functions keep the reference's signatures (names, order, keyword-only split, defaults).

The reference sources are PARSED (``ast``), not imported — ``scanpy`` / ``anndata`` are not installed — so this also
covers ``tl.pca`` / ``tl.leiden`` / ``pp.neighbors``.  Needs ``/root/reference`` (build container); skipped elsewhere.
"""

import ast
import inspect
from pathlib import Path

import pytest

import infercnvpy_b200 as cnv

REF = Path("/root/reference/src/infercnvpy")

CASES = [
    ("tl/_infercnv.py", "infercnv", cnv.tl.infercnv),
    ("tl/_scores.py", "cnv_score", cnv.tl.cnv_score),
    ("tl/_scores.py", "ithgex", cnv.tl.ithgex),
    ("tl/_scores.py", "ithcna", cnv.tl.ithcna),
    ("tl/__init__.py", "pca", cnv.tl.pca),
    ("tl/__init__.py", "leiden", cnv.tl.leiden),
    ("pp/__init__.py", "neighbors", cnv.pp.neighbors),
    ("tl/__init__.py", "umap", cnv.tl.umap),
    ("tl/__init__.py", "tsne", cnv.tl.tsne),
]


def _ref_signature(path: Path, name: str):
    tree = ast.parse(path.read_text())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    pos_defaults = [None] * (len(pos) - len(a.defaults)) + [ast.literal_eval(d) for d in a.defaults]
    kwonly = [x.arg for x in a.kwonlyargs]
    kw_defaults = [None if d is None else ast.literal_eval(d) for d in a.kw_defaults]
    return pos, pos_defaults, kwonly, kw_defaults, a.vararg is not None, a.kwarg is not None


@pytest.mark.skipif(not REF.is_dir(), reason="reference sources not present on this machine")
@pytest.mark.parametrize("rel,name,ours", CASES, ids=[c[1] for c in CASES])
def test_signature_matches_reference(rel, name, ours):
    pos, pos_defaults, kwonly, kw_defaults, has_var, has_kw = _ref_signature(REF / rel, name)
    sig = inspect.signature(ours)
    P = inspect.Parameter
    ours_pos = [p for p in sig.parameters.values() if p.kind in (P.POSITIONAL_ONLY, P.POSITIONAL_OR_KEYWORD)]
    ours_kw = [p for p in sig.parameters.values() if p.kind == P.KEYWORD_ONLY]
    assert [p.name for p in ours_pos] == pos
    assert [p.name for p in ours_kw] == kwonly
    for p, d in zip(ours_pos, pos_defaults):
        assert (None if p.default is P.empty else p.default) == d, p.name
    for p, d in zip(ours_kw, kw_defaults):
        got = None if p.default is P.empty else p.default
        assert got == d, p.name
    assert any(p.kind == P.VAR_POSITIONAL for p in sig.parameters.values()) == has_var
    assert any(p.kind == P.VAR_KEYWORD for p in sig.parameters.values()) == has_kw
