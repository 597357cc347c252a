"""Deterministic, platform-independent weights for parity tests (TEST INFRASTRUCTURE).

The golden vectors under tests/golden/ were produced by loading exactly these tensors into the reference
modules (oracle/make_golden.py).  They are regenerated from a {key: shape} spec with numpy's legacy
RandomState (bit-stable by numpy policy), so 275 MB of checkpoints never needs to be committed.
"""
import zlib

import numpy as np
import torch


def _rs(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def make_tensor(key, shape, seed=0):
    """Scales follow the reference's initialisers so that the test network behaves like a freshly constructed one:
    nn.Conv2d / nn.Linear default (uniform +-1/sqrt(fan_in) for weight and bias), kaiming_normal(fan_out) with zero-ish
    bias for the VGG encoder (vgg_tro_channel3_modi.py:29-37), N(0, 1) embeddings."""
    rs = _rs(key, seed)
    shape = tuple(int(s) for s in shape)
    leaf = key.rsplit(".", 1)[-1]
    f32 = np.float32
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return torch.from_numpy((0.1 * rs.standard_normal(shape)).astype(f32))
    if leaf == "running_var":
        return torch.from_numpy((1.0 + 0.1 * np.abs(rs.standard_normal(shape))).astype(f32))
    if len(shape) >= 2:
        if "embed" in key:
            return torch.from_numpy(rs.standard_normal(shape).astype(f32))
        if "enc_image" in key:
            fan_out = shape[0] * int(np.prod(shape[2:]))
            return torch.from_numpy(((2.0 / fan_out) ** 0.5 * rs.standard_normal(shape)).astype(f32))
        bound = 1.0 / int(np.prod(shape[1:])) ** 0.5
        return torch.from_numpy(rs.uniform(-bound, bound, shape).astype(f32))
    # 1-D tensors: biases and norm scales (the caller shifts norm scales to be centred on 1)
    return torch.from_numpy((0.05 * rs.standard_normal(shape)).astype(f32))


def make_state(spec, seed=0):
    """spec: {key: shape}.  1-D `weight` tensors that belong to a norm layer (a sibling running_mean exists)
    are drawn around 1, every other 1-D tensor around 0."""
    out = {}
    for key, shape in spec.items():
        t = make_tensor(key, shape, seed)
        if key.endswith(".weight") and len(shape) == 1 and key[:-len("weight")] + "running_mean" in spec:
            t = t + 1.0
        out[key] = t
    return out


def spec_of(module_or_state):
    sd = module_or_state.state_dict() if hasattr(module_or_state, "state_dict") else module_or_state
    return {k: list(v.shape) for k, v in sd.items()}


def alias_extractor(sd):
    """torchvision-ResNet encoders: `extractor.*` (the fx GraphModule) shares its tensors with `model.*` in the reference
    (modules_tro.py:503), so a state built key by key must carry identical values under both prefixes."""
    for k in list(sd):
        if k.startswith("extractor.") and "model." + k[len("extractor."):] in sd:
            sd[k] = sd["model." + k[len("extractor."):]]
    return sd
