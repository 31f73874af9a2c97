"""Import the reference's OWN torch networks by path (test infrastructure only).

This module is ORACLE code: only tests/, bench.py's cpu_baseline / --impl reference
leg and oracle/make_golden.py may import it.  It needs /root/reference and is
therefore usable only in the build container, never on the GPU box.

Recipe: SURVEY.md Appendix C.  Mirrors rapid_doc/model/ocr/torch.py:69-116
(TorchInferSession._build_and_load_model): build BaseModel from
rapid_doc/resources/arch_config.yaml, strip the "model." key prefix, take
out_channels from head.head.weight.
"""
import copy
import os
import sys
import types

REF = os.environ.get("RAPIDDOC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "rapid_doc", "model", "ocr", "ppocrv6_pytorch"))


def _stub(name, path):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m


def _import_base_model():
    # avoid rapid_doc/__init__.py (imports pypdfium2 through main.py)
    _stub("rapid_doc", f"{REF}/rapid_doc")
    for sub in ["model", "model.ocr", "model.ocr.ppocrv6_pytorch", "model.ocr.ppocrv6_pytorch.modeling"]:
        _stub("rapid_doc." + sub, f"{REF}/rapid_doc/" + sub.replace(".", "/"))
    from rapid_doc.model.ocr.ppocrv6_pytorch.modeling.architectures.base_model import BaseModel
    return BaseModel


def build(arch_key: str, weights_file: str):
    import yaml
    from safetensors.torch import load_file
    BaseModel = _import_base_model()
    arch = yaml.safe_load(open(f"{REF}/rapid_doc/resources/arch_config.yaml"))
    sd = {k.removeprefix("model."): v for k, v in load_file(f"{REF}/rapid_doc/resources/{weights_file}").items()}
    kw = {"out_channels": int(sd["head.head.weight"].shape[0])} if "head.head.weight" in sd else {}
    net = BaseModel(copy.deepcopy(arch[arch_key]), **kw)
    net.load_state_dict(sd)
    return net.eval()


def det_net():
    """x[B,3,H,W] f32 -> {"maps": [B,1,H,W]} (reference engine output: torch.py:181-184)."""
    return build("ch_PP-OCRv6_det_small", "ch_PP-OCRv6_det_small.safetensors")


def rec_net():
    """x[B,3,48,W] f32 -> {"ctc_logits": [B,W/8,18710]}; engine applies softmax (torch.py:186-187)."""
    return build("ch_PP-OCRv6_small_rec_infer", "ch_PP-OCRv6_rec_small.safetensors")


def characters():
    chars = ["blank"] + [l.rstrip("\n") for l in open(f"{REF}/rapid_doc/resources/ppocrv6_small_dict.txt", encoding="utf-8")] + [" "]
    assert len(chars) == 18710
    return chars


def formula_net(max_new_tokens=16):
    """The reference's PP-FormulaNet_plus-M torch module (random init; load a state_dict into it), imported by path:
    rapid_doc/model/formula/rapid_formula_self/networks (arch config pp_formulanet_arch_config.yaml).
    x [B,1,384,384] f32 -> ids [B,L] int64 (BaseModel.forward -> PPFormulaNet_Head.generate_export in eval mode)."""
    import os
    import yaml
    os.environ["RAPID_FORMULA_DEVICE_MODE"] = "cpu"
    if "rapid_doc" not in sys.modules:
        _stub("rapid_doc", f"{REF}/rapid_doc")
    for sub in ["model", "model.formula", "model.formula.rapid_formula_self", "model.formula.rapid_formula_self.networks"]:
        if "rapid_doc." + sub not in sys.modules:
            _stub("rapid_doc." + sub, f"{REF}/rapid_doc/" + sub.replace(".", "/"))
    from rapid_doc.model.formula.rapid_formula_self.networks.architectures.base_model import BaseModel
    cfg = yaml.safe_load(open(f"{REF}/rapid_doc/model/formula/rapid_formula_self/networks/pp_formulanet_arch_config.yaml"))["PP-FormulaNet_plus-M"]
    cfg = copy.deepcopy(cfg)
    cfg["Head"]["max_new_tokens"] = int(max_new_tokens)
    return BaseModel(cfg).eval()
