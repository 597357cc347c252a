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


# ------------------------------------------------------------------------------------------------ uint8 wire format
def decode_u8(u8, out=None):
    """Device-side half of `IAM_words.read_image_single` (load_data.py:152-166): grey-level uint8 pixels as cv2 leaves them
    after the resize (right padding = 255) -> the normalised float32 image the models take, bit for bit
    (`affgw_u8_to_image`).  Shipping the uint8 canvases and normalising on the GPU moves a quarter of the bytes over PCIe
    (SURVEY.md §8(f).3).  u8: CUDA uint8 tensor of any shape; returns a float32 tensor of the same shape."""
    import torch
    from . import _lib as L
    L.require_cuda(u8)
    if u8.dtype != torch.uint8:
        raise RuntimeError("decode_u8: expected a uint8 tensor, got %s" % u8.dtype)
    u8 = u8.contiguous()
    if out is None:
        out = torch.empty(u8.shape, dtype=torch.float32, device=u8.device)
    elif out.shape != u8.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != u8.device:
        raise RuntimeError("decode_u8: `out` must be a contiguous float32 tensor of the input's shape and device")
    if u8.numel():
        L.call("affgw_u8_to_image", u8.data_ptr(), out.data_ptr(), u8.numel(), L.stream())
    return out


def batch_to_device(batch, device):
    """Host batch (the 9-tuple of main_run.py:108-118, images either float32 as the reference's DataLoader yields them or
    uint8 wire format) -> device batch: uint8 tensors are copied as bytes and normalised on the GPU."""
    import torch
    out = []
    for t in batch:
        if torch.is_tensor(t):
            t = t.to(device, non_blocking=True)
            if t.dtype == torch.uint8:
                t = decode_u8(t)
        out.append(t)
    return tuple(out)


class DevicePrefetcher:
    """Double-buffered host -> device staging of batches on a copy stream, so that the PCIe transfer and the uint8
    normalisation of batch k+1 overlap the training step of batch k (the role of the reference's pinned DataLoader workers,
    main_run.py:52,123-130, on the device side).

        pf.stage(host_batch)            # asynchronous: H2D (+ affgw_u8_to_image) into the next slot, on the copy stream
        batch = pf.get()                # the oldest staged batch; the current stream waits for its copy
        ... train_step(batch) ...
        pf.release()                    # the slot may be overwritten once the work queued so far has run

    Slots are allocated once (the first batch that visits a slot fixes its shapes), nothing is allocated per step."""

    def __init__(self, device, slots=2):
        import torch
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [None] * slots
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [None] * slots
        self._head = 0          # next slot to stage into
        self._tail = 0          # next slot to hand out
        self._out = []          # slots handed out and not released yet
        self._staged = {}       # slot -> device batch

    def _buffers(self, i, host_batch):
        import torch
        if self.slots[i] is None:
            bufs = []
            for t in host_batch:
                if not torch.is_tensor(t):
                    bufs.append(None)
                elif t.dtype == torch.uint8:
                    bufs.append((torch.empty(t.shape, dtype=torch.uint8, device=self.device),
                                 torch.empty(t.shape, dtype=torch.float32, device=self.device)))
                else:
                    bufs.append((torch.empty(t.shape, dtype=t.dtype, device=self.device), None))
            self.slots[i] = bufs
        return self.slots[i]

    def stage(self, host_batch):
        import torch
        i = self._head % len(self.slots)
        if self._head - self._tail >= len(self.slots) or i in self._out:
            raise RuntimeError("DevicePrefetcher: every slot is staged or in use (call get() / release() first)")
        bufs = self._buffers(i, host_batch)
        if self.free[i] is not None:
            self.stream.wait_event(self.free[i])
        staged = []
        with torch.cuda.stream(self.stream):
            for t, b in zip(host_batch, bufs):
                if b is None:
                    staged.append(t)
                    continue
                raw, img = b
                if raw.shape != t.shape or raw.dtype != t.dtype:
                    raise RuntimeError("DevicePrefetcher: batch layout changed (%s %s, slot holds %s %s)"
                                       % (tuple(t.shape), t.dtype, tuple(raw.shape), raw.dtype))
                raw.copy_(t, non_blocking=True)
                staged.append(decode_u8(raw, out=img) if img is not None else raw)
            self.ready[i].record(self.stream)
        self._staged[i] = tuple(staged)
        self._head += 1

    def get(self):
        import torch
        if self._tail == self._head:
            raise RuntimeError("DevicePrefetcher: nothing staged")
        i = self._tail % len(self.slots)
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self._tail += 1
        self._out.append(i)
        return self._staged[i]

    def release(self):
        import torch
        i = self._out.pop(0)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[i] = ev
