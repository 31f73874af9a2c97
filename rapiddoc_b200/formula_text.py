"""Token ids -> LaTeX string for PP-FormulaNet, restated without an importable `rapid_doc`:

  FormulaDecode           `UniMERNetDecode` token2str / post_process / fix_latex / remove_chinese_text_wrapping / normalize
                          rapid_formula_self/model_handler/pp_formulanet_plus/post_process.py:272-389
  fix_latex_left_right, fix_left_right_pairs, fix_latex_environments, remove_up_commands, remove_unsupported_commands
                          .../pp_formulanet_plus/utils.py:9-146,253-313

Pure host string work (a few hundred characters per formula).  The tokenizer (`tokenizer.json`, shipped with the checkpoint) is
a Hugging Face `tokenizers.Tokenizer` built by the caller — `FormulaDecode(tokenizer)` only needs its `decode(ids,
skip_special_tokens=True)`.  `ftfy.fix_text` (the last step of the reference's post_process) is applied when the package is
importable and skipped otherwise (`FormulaDecode.ftfy_applied` says which).  tests/test_formula_text.py runs every function
against the reference's own module imported by path.
"""
import re

import numpy as np

_VALID_DELIMS = {"(", ")", "[", "]", "{", "}", "/", "|", r"\{", r"\}", r"\lceil", r"\rceil", r"\lfloor", r"\rfloor", r"\backslash",
                 r"\uparrow", r"\downarrow", r"\Uparrow", r"\Downarrow", r"\|", r"\."}
_LEFT, _RIGHT = re.compile(r"(\\left)(\S*)"), re.compile(r"(\\right)(\S*)")
_LEFT_N, _RIGHT_N = re.compile(r"\\left(?![a-zA-Z])"), re.compile(r"\\right(?![a-zA-Z])")
_LR_REMOVE = re.compile(r"\\left\.?|\\right\.?")
ENV_TYPES = ["array", "matrix", "pmatrix", "bmatrix", "vmatrix", "Bmatrix", "Vmatrix", "cases", "aligned", "gathered", "align", "align*"]


def _escaped(text, pos):
    """An odd number of backslashes right before text[pos]."""
    n, j = 0, pos - 1
    while j >= 0 and text[j] == "\\":
        n += 1
        j -= 1
    return n % 2 == 1


def _group_end(text, pos, depth):
    """Index of the `}` that closes the brace group of nesting `depth` containing `pos` (-1: none)."""
    cur = depth
    for i in range(pos, len(text)):
        if text[i] == "{" and (i == 0 or not _escaped(text, i)):
            cur += 1
        elif text[i] == "}" and (i == 0 or not _escaped(text, i)):
            cur -= 1
            if cur < depth:
                return i
    return -1


def fix_left_right_pairs(s):
    """A `\\right` that closes a `\\left` opened at another brace depth is moved to the end of the `\\left`'s group."""
    braces, lefts, moves = [], [], []
    i, n = 0, len(s)
    while i < n:
        if i > 0 and s[i - 1] == "\\" and _escaped(s, i):
            i += 1
            continue
        if i + 5 < n and s[i:i + 5] == "\\left":
            lefts.append((i, len(braces)))
            i += 6
            continue
        if i + 6 < n and s[i:i + 6] == "\\right":
            if lefts:
                lpos, ldepth = lefts.pop()
                if ldepth != len(braces):
                    t = _group_end(s, lpos, ldepth)
                    if t != -1:
                        moves.append((i, i + 7, t))
            i += 7
            continue
        if s[i] == "{":
            braces.append(i)
        elif s[i] == "}" and braces:
            braces.pop()
        i += 1
    if not moves:
        return s
    out = list(s)
    for a, b, t in sorted(moves, key=lambda m: m[0], reverse=True):
        piece = out[a:b]
        del out[a:b]
        out.insert(t, "".join(piece))
    return "".join(out)


def fix_latex_left_right(s, fix_delimiter=True):
    """`\\left` / `\\right` get a `.` when no valid delimiter follows; unbalanced counts -> all of them removed; balanced ->
    pairs split across brace groups are repaired."""
    if fix_delimiter:
        def fix(m):
            return m.group(1) + "." if not m.group(2) or m.group(2) not in _VALID_DELIMS else m.group(0)
        s = _RIGHT.sub(fix, _LEFT.sub(fix, s))
    if len(_LEFT_N.findall(s)) == len(_RIGHT_N.findall(s)):
        return fix_left_right_pairs(s)
    return _LR_REMOVE.sub("", s)


def fix_latex_environments(s):
    """Missing `\\begin{env}` are prepended (with the format of the first one found, `{c}` for array), missing `\\end{env}` appended.
    (The environment name goes into the pattern unescaped, as in the reference: `align*` therefore counts `alig`, `align`, ...)"""
    for env in ENV_TYPES:
        nb = len(re.findall(r"\\begin\{" + env + r"\}", s))
        ne = len(re.findall(r"\\end\{" + env + r"\}", s))
        if nb == ne:
            continue
        if ne > nb:
            m = re.search(r"\\begin\{" + env + r"\}\{([^}]*)\}", s)
            fmt = "{" + m.group(1) + "}" if m else ("{c}" if env == "array" else "")
            s = ("\\begin{" + env + "}" + fmt + " ") * (ne - nb) + s
        else:
            s = s + (" \\end{" + env + "}") * (nb - ne)
    return s


def remove_up_commands(s):
    return re.sub(r"\\up([a-zA-Z]+)", lambda m: m.group(0) if m.group(1) in ("arrow", "downarrow", "lus", "silon") else "\\" + m.group(1), s)


def remove_unsupported_commands(s):
    return re.sub(r"\\(?:lefteqn|boldmath|ensuremath|centering|textsubscript|sides|textsl|textcent|emph|protect|null)", "", s)


def remove_chinese_text_wrapping(formula):
    """`\\text{...CJK...}` -> its content; double quotes dropped."""
    return re.sub(r"\\text\s*{\s*([^}]*?[\u4e00-\u9fff]+[^}]*?)\s*}", lambda m: m.group(1), formula).replace('"', "")


def normalize(s):
    """Whitespace removal between non-letters / letters as UniMERNetDecode.normalize does it (kept for parity: the reference's
    post_process has the call commented out)."""
    text_reg = r"(\\(operatorname|mathrm|text|mathbf)\s?\*? {.*?})"
    letter, noletter = "[a-zA-Z]", r"[\W_^\d]"
    names = []
    for x in re.findall(text_reg, s):
        for m in re.findall(r"(\\[a-zA-Z]+)\s(?=\w)|\\[a-zA-Z]+\s(?=})", x[0]):
            if m not in ("\\operatorname", "\\mathrm", "\\text", "\\mathbf") and m.strip() != "":
                s = s.replace(m, m + "XXXXXXX").replace(" ", "")
                names.append(s)
    if names:
        s = re.sub(text_reg, lambda match: str(names.pop(0)), s)
    news = s
    while True:
        s = news
        news = re.sub(r"(?!\\ )(%s)\s+?(%s)" % (noletter, noletter), r"\1\2", s)
        news = re.sub(r"(?!\\ )(%s)\s+?(%s)" % (noletter, letter), r"\1\2", news)
        news = re.sub(r"(%s)\s+?(%s)" % (letter, noletter), r"\1\2", news)
        if news == s:
            break
    return s.replace("XXXXXXX", " ")


def fix_latex(text):
    return remove_unsupported_commands(remove_up_commands(fix_latex_environments(fix_latex_left_right(text, fix_delimiter=False))))


class FormulaDecode:
    """ids [B, T] -> list of LaTeX strings: cut at the first eos (id 2, kept), tokenizer.decode(skip_special_tokens=True),
    remove_chinese_text_wrapping, fix_latex, ftfy.fix_text."""
    eos_token_id = 2

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer
        try:
            from ftfy import fix_text
            self._fix_text, self.ftfy_applied = fix_text, True
        except ImportError:
            self._fix_text, self.ftfy_applied = (lambda t: t), False

    def post_process(self, text):
        return self._fix_text(fix_latex(remove_chinese_text_wrapping(text)))

    def token2str(self, token_ids):
        out = []
        for row in token_ids:
            row = np.asarray(row)
            end = np.argwhere(row == self.eos_token_id)
            if len(end) > 0:
                row = row[: int(end[0][0]) + 1]
            out.append(self.post_process(self.tokenizer.decode([int(t) for t in row], skip_special_tokens=True)))
        return out

    def __call__(self, preds, label=None, mode="eval"):
        preds = np.array(preds)
        text = self.token2str(preds.argmax(axis=2) if mode == "train" else preds)
        return text if label is None else (text, self.token2str(np.array(label)))

    def decode_row(self, ids):
        """`decode=` hook of B200FormulaModel: one row of ids -> string."""
        return self.token2str([ids])[0]
