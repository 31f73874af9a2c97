"""F5: the ids -> LaTeX text stage against the reference's own functions (imported by path) on adversarial strings."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rapiddoc_b200 import formula_text as FT      # noqa: E402

REF_DIR = "/root/reference/rapid_doc/model/formula/rapid_formula_self/model_handler/pp_formulanet_plus"

CASES = [
    r"\left( a + b \right)",
    r"\left a + \right b",
    r"\left( \frac { a } { b \right) }",
    r"\frac { \left( a } { b } \right)",
    r"\left( x \right) \left[ y",
    r"\leftarrow x \rightarrow y \left. z \right|",
    r"\left\{ \begin{array}{cc} a & b \\ c & d \end{array} \right.",
    r"a & b \\ c & d \end{array}",
    r"\begin{array}{ll} a & b \\ c & d \end{array} x \end{array}",
    r"\begin{cases} x \\ y",
    r"\begin{align*} x \end{align*} \end{align*}",
    r"\begin{matrix} 1 \end{matrix} \end{matrix}",
    r"\upalpha + \uparrow + \updownarrow + \uplus + \upsilon + \upbeta",
    r"\lefteqn { x } \boldmath \emph { y } \protect \null z \textsl a",
    r'\text { 中文 } + \text { abc } + \text{速度 v} "q"',
    r"\\left( a \\right)",
    r"\{ \left( a \} \right)",
    r"x ^ { 2 } + \mathrm { d } x \, \operatorname { sin } \theta",
    r"\mathrm { \alpha b } + \text { a b } c _ { 1 2 }",
    "",
    r"\left",
    r"\right)",
    r"{ { \left( } a \right) }",
    r"\left( { a \right) } \left[ { b } \right]",
]


def _ref_utils():
    if not os.path.isdir(REF_DIR):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location("ref_ppf_utils", os.path.join(REF_DIR, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_decode_class(monkeypatch):
    """post_process.py with its package-absolute import satisfied by the utils module loaded by path."""
    utils = _ref_utils()
    names = ["rapid_doc", "rapid_doc.model", "rapid_doc.model.formula", "rapid_doc.model.formula.rapid_formula_self",
             "rapid_doc.model.formula.rapid_formula_self.model_handler", "rapid_doc.model.formula.rapid_formula_self.model_handler.pp_formulanet_plus"]
    for n in names:
        m = types.ModuleType(n)
        m.__path__ = []
        monkeypatch.setitem(sys.modules, n, m)
    monkeypatch.setitem(sys.modules, names[-1] + ".utils", utils)
    spec = importlib.util.spec_from_file_location("ref_ppf_post", os.path.join(REF_DIR, "post_process.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.UniMERNetDecode


@pytest.mark.parametrize("name", ["fix_latex_left_right", "fix_left_right_pairs", "fix_latex_environments", "remove_up_commands", "remove_unsupported_commands"])
def test_latex_fix_functions_equal_the_reference(name):
    ref = _ref_utils()
    rng = np.random.RandomState(0)
    toks = [r"\left", r"\right", "(", ")", "[", "]", "{", "}", r"\{", r"\}", ".", "|", " ", "a", r"\\", r"\frac", r"\begin{array}", r"{cc}", r"\end{array}",
            r"\begin{cases}", r"\end{cases}", r"\upalpha", r"\emph", r"\leftarrow", "x", "^", "_", "&"]
    fuzz = ["".join(toks[i] + (" " if rng.rand() < 0.5 else "") for i in rng.randint(0, len(toks), rng.randint(1, 30))) for _ in range(400)]
    for s in CASES + fuzz:
        assert getattr(FT, name)(s) == getattr(ref, name)(s), s
    if name == "fix_latex_left_right":
        for s in CASES + fuzz[:100]:
            assert FT.fix_latex_left_right(s, fix_delimiter=False) == ref.fix_latex_left_right(s, fix_delimiter=False), s


def test_decode_class_equals_the_reference(monkeypatch):
    pytest.importorskip("tokenizers")
    Ref = _ref_decode_class(monkeypatch)
    ref = Ref.__new__(Ref)                                   # the string methods need no tokenizer state
    for s in CASES:
        assert FT.remove_chinese_text_wrapping(s) == ref.remove_chinese_text_wrapping(s), s
        assert FT.normalize(s) == ref.normalize(s), s
        assert FT.fix_latex(s) == ref.fix_latex(s), s

    class Tok:                                               # stand-in for tokenizers.Tokenizer: ids index a table of strings
        table = ["<s>", "<pad>", "</s>"] + [c + " " for c in CASES]

        def decode(self, ids, skip_special_tokens=True):
            return "".join(self.table[i] for i in ids if not (skip_special_tokens and i < 3)).strip()
    ids = np.array([[0, 5, 9, 2, 1, 1], [0, 3, 2, 7, 1, 1], [0, 17, 10, 11, 12, 13]])
    mine = FT.FormulaDecode(Tok())
    ref.tokenizer = Tok()
    if not mine.ftfy_applied:                                # ftfy absent: the reference's post_process cannot run; compare its stages
        ref.post_process = lambda t: ref.fix_latex(ref.remove_chinese_text_wrapping(t))
    assert mine(ids) == ref(ids)
    assert mine.decode_row(ids[1]) == mine(ids)[1] and "7" not in mine(ids)[1]          # cut at the first eos


def test_decode_hook_plugs_into_the_formula_model():
    class Tok:
        def decode(self, ids, skip_special_tokens=True):
            return " ".join(f"t{i}" for i in ids if i > 2)
    d = FT.FormulaDecode(Tok())
    assert d.decode_row(np.array([0, 7, 8, 2, 9])) == "t7 t8"
    assert FT.fix_latex(r"\left( a") == "( a" and FT.fix_latex(r"x \end{cases}") == r"\begin{cases} x \end{cases}"
