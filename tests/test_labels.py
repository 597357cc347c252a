"""CPU: integer label handling is bit-exact with the reference (known-answer vectors
in tests/golden/labels.json were produced by the reference's own IAM_words.label_padding and TextEncoder_FC)."""
import json
import os

import pytest

from affganwriting_b200 import load_data
from oracle import affgw_oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "labels.json")))


def test_known_answer_three():
    assert load_data.label_padding("three") == [0, 22, 10, 20, 7, 7, 1, 2, 2, 2, 2, 2]


@pytest.mark.parametrize("word", sorted(GOLD["label_padding"]))
def test_label_padding_matches_reference(word):
    ref = GOLD["label_padding"][word]
    assert load_data.label_padding(word) == ref
    assert O.label_padding(word) == ref


def test_vocabulary_constants():
    assert load_data.vocab_size == GOLD["vocab_size"] == O.VOCAB_SIZE == 55
    assert load_data.tokens == GOLD["tokens"] == O.TOKENS
    assert load_data.OUTPUT_MAX_LEN == 12 and load_data.IMG_HEIGHT == 64 and load_data.IMG_WIDTH == 216
    assert load_data.letter2index["a"] == 0 and load_data.letter2index["Z"] == 51


def test_maximum_length_and_overflow():
    assert load_data.label_padding("abcdefghij")[-1] == 1          # 10 chars: GO + 10 + END, no PAD
    with pytest.raises(ValueError):
        load_data.label_padding("abcdefghijk")
    with pytest.raises(KeyError):
        load_data.label_padding("a-b")                             # outside a-zA-Z, as in the reference


@pytest.mark.parametrize("width", sorted(GOLD["text_column_map"], key=int))
def test_text_column_map(width):
    """column -> token slot of the content map (modules_tro.py:295-313); the CUDA kernel uses the same rule
    (slot = c // reps for c < ts * reps, PAD otherwise)."""
    ref = GOLD["text_column_map"][width]
    w, ts = int(width), 12
    assert O.text_column_map(w) == ref
    reps = max(1, w // ts)
    mine = [(c // reps if c < ts * reps else -1) for c in range(ts * reps + w % ts)]
    assert mine == ref
