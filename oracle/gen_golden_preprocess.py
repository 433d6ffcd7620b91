"""ORACLE tooling: golden vectors for the OCR-annotation preprocessor, produced by the REFERENCE's own
``pixparse.data.preprocess.preprocess_ocr_anno`` / ``preprocess_text_anno`` (data/preprocess.py:9-110) imported from
/root/reference over oracle/ref_shims.py. Runs only in the build container.

    python -m oracle.gen_golden_preprocess        # writes tests/golden/preprocess_anno.json
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from pixparse_b200 import synthetic  # noqa: E402


def main():
    ref_shims.install()
    from pixparse.data.preprocess import preprocess_ocr_anno, preprocess_text_anno
    cases = []
    for seed in range(12):
        tok = synthetic.CharTokenizer()
        tok.add_special_tokens({"additional_special_tokens": sorted({"<sep/>", "<s_pretrain>"})})
        anno = synthetic.synthetic_ocr_annotation(seed)
        if seed == 7:
            anno = [17, anno]            # the legacy [id, {...}] form
        out, info = preprocess_ocr_anno(anno, tok, 48, "<s_pretrain>", "<s_pretrain>", generator=random.Random(100 + seed))
        cases.append({"seed": seed, "kind": "ocr", "text": [t.tolist() for t in out["text"]],
                      "target": [t.tolist() for t in out["target"]], "info": info})
    for seed in range(3):
        tok = synthetic.CharTokenizer()
        tok.add_special_tokens({"additional_special_tokens": sorted({"<sep/>", "<s_pretrain>"})})
        raw = "line %d of raw text " % seed * (seed + 1)
        out = preprocess_text_anno(raw, tok, 40, "<s_pretrain>", "<s_pretrain>")
        cases.append({"seed": seed, "kind": "text", "raw": raw, "text": [t.tolist() for t in out["text"]],
                      "target": [t.tolist() for t in out["target"]]})
    path = os.path.join(ROOT, "tests", "golden", "preprocess_anno.json")
    with open(path, "w") as f:
        json.dump({"source": "pixparse.data.preprocess (reference, unmodified)", "max_len_ocr": 48, "max_len_text": 40,
                   "cases": cases}, f)
    print(path, len(cases))


if __name__ == "__main__":
    main()
