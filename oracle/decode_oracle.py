"""CPU restatement of the reference's token -> SMILES decode.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``reverse_tokenize`` (MoleculeDiffusion/generative.py:1069-1078):

    X_data_tokenized_reversed = tokenizer_X.sequences_to_texts((X_data * X_norm_factor).astype(int))
    ... str(i).replace(' ', '') for every row

``tokenizer_X`` is a Keras ``Tokenizer`` (tensorflow.keras.preprocessing.text, not installed here and not vendored by the
reference; Keras 2.x ``Tokenizer.sequences_to_texts_generator``): for every id of a row it looks the id up in ``index_word``;
ids without an entry are skipped unless an ``oov_token`` was configured (the notebooks configure none), and the words found are
joined with ' '.  Keras itself is not installed here, so the pin is the reference's own recorded evidence: the tokenizer
vocabulary, six (token row, ``reverse_tokenize`` output) pairs and eight decoded SMILES strings that the authors' run of
``Inverse_Diffusion.ipynb`` left in its cell outputs (cells 36, 38, 65), extracted by ``oracle/make_decode_fixture.py`` into
``tests/golden/decode_notebook.json`` and checked in ``tests/test_host_cpu.py::test_decode_oracle_against_notebook_record``.
"""
from typing import Dict, List, Sequence


def sequences_to_texts(index_word: Dict[int, str], sequences: Sequence[Sequence[int]]) -> List[str]:
    texts = []
    for seq in sequences:
        words = []
        for num in seq:
            word = index_word.get(int(num))
            if word is not None:
                words.append(word)
        texts.append(" ".join(words))
    return texts


def reverse_tokenize(index_word: Dict[int, str], x_data, x_norm_factor=1) -> List[str]:
    """generative.py:1069-1078 with ``tokenizer_X.index_word`` passed directly."""
    import numpy as np

    ids = (np.asarray(x_data) * x_norm_factor).astype(int)
    return [str(t).replace(" ", "") for t in sequences_to_texts(index_word, ids)]
