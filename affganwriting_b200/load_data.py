"""Constants and label handling of the reference (GAN_word/load_data.py:11-19, 31-40, 169-179).

Only the pieces the generator / discriminator path needs: no dataset files are opened at import.
Integer label handling is bit-exact with the reference (tests/test_labels.py pins it against vectors the
reference itself produced).
"""
import string

IMG_HEIGHT = 64
IMG_WIDTH = 216
MAX_CHARS = 10
NUM_CHANNEL = 50          # stacked style images per writer (15 in the original GANwriting)
EXTRA_CHANNEL = NUM_CHANNEL + 1
NUM_WRITERS = 500
NORMAL = True
OUTPUT_MAX_LEN = MAX_CHARS + 2  # <GO> + groundtruth + <END>


def labelDictionary():
    labels = list(string.ascii_lowercase + string.ascii_uppercase)
    letter2index = {label: n for n, label in enumerate(labels)}
    index2letter = {v: k for k, v in letter2index.items()}
    return len(labels), letter2index, index2letter


num_classes, letter2index, index2letter = labelDictionary()
tokens = {"GO_TOKEN": 0, "END_TOKEN": 1, "PAD_TOKEN": 2}
num_tokens = len(tokens.keys())
vocab_size = num_classes + num_tokens


def label_padding(labels, num_tokens=num_tokens, output_max_len=OUTPUT_MAX_LEN):
    """IAM_words.label_padding: chars -> letter2index + num_tokens, GO first, END last, PAD to output_max_len."""
    ll = [letter2index[i] + num_tokens for i in labels]
    ll = [tokens["GO_TOKEN"]] + ll + [tokens["END_TOKEN"]]
    num = output_max_len - len(ll)
    if num < 0:
        raise ValueError(f"word {labels!r} longer than {output_max_len - 2} characters")
    if not num == 0:
        ll.extend([tokens["PAD_TOKEN"]] * num)
    return ll
