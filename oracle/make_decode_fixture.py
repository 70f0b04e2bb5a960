"""Extract the reference-held known answers for the token -> SMILES decode from the reference's own notebook.

    python -m oracle.make_decode_fixture        # writes tests/golden/decode_notebook.json

``Inverse_Diffusion.ipynb`` records (a) the fitted Keras tokenizer's ``index_word`` / ``word_index`` (cell 36 output), (b) six
tokenised, zero-padded rows next to what the reference's ``reverse_tokenize`` (generative.py:1069-1078) returned for them
(cell 38 output) and (c) SMILES strings the reference decoded from sampled tokens (cell 65 output: "Result as SMILES" / "GT as
SMILES").  (a)+(b) are direct input/output pairs of the function the oracle restates; (c) are strings that must survive
tokenise -> decode with the recorded vocabulary.  Only these few values are stored, not the notebook.
"""
from __future__ import annotations

import ast
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOTEBOOK = os.path.join(os.environ.get("MDT_REFERENCE_ROOT", "/root/reference"), "Inverse_Diffusion.ipynb")


def _outputs(cell):
    for o in cell.get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain")
        if t:
            yield "".join(t)


def main():
    cells = json.load(open(NOTEBOOK))["cells"]
    # (a) tokenizer config printed by cell 36
    cfg_text = next(t for t in _outputs(cells[36]) if "index_word" in t)
    cfg = ast.literal_eval(cfg_text[cfg_text.index("{"): cfg_text.rindex("}") + 1])
    index_word = {int(k): v for k, v in json.loads(cfg["index_word"]).items()}
    word_index = json.loads(cfg["word_index"])
    assert cfg["char_level"] and cfg["oov_token"] is None and cfg["filters"] == ""
    # (b) tokenised rows and their decoded strings, cell 38
    out38 = next(_outputs(cells[38]))
    arr = re.search(r"array\((\[\[.*?\]\])", out38, re.S).group(1)
    rows = ast.literal_eval(re.sub(r"\s+", " ", arr))
    decoded = ast.literal_eval(re.search(r"\[('[^\]]*')\]\)\s*$", out38.strip(), re.S).group(0)[:-1])
    assert len(rows) == len(decoded) == 6
    # (c) decoded SMILES printed by the sampling loop, cell 65
    text65 = "\n".join(_outputs(cells[65]))
    smiles = []
    for key in ("Result as SMILES:", "GT as SMILES:"):
        m = re.search(re.escape(key) + r"\s*(\[.*?\])", text65)
        smiles += ast.literal_eval(m.group(1))
    x_norm_factor = int(next(_outputs(cells[40])).strip())
    fixture = {"source": "Inverse_Diffusion.ipynb cells 36, 38, 40, 65 (outputs recorded by the reference's authors)",
               "index_word": {str(k): v for k, v in index_word.items()}, "word_index": word_index, "x_norm_factor": x_norm_factor,
               "tokenized_rows": rows, "reverse_tokenized": decoded, "decoded_smiles": smiles}
    path = os.path.join(ROOT, "tests", "golden", "decode_notebook.json")
    json.dump(fixture, open(path, "w"), indent=1)
    print(f"wrote {path}: {len(index_word)} vocabulary entries, {len(rows)} row pairs, {len(smiles)} decoded strings")


if __name__ == "__main__":
    main()
