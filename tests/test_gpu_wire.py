"""uint8 wire format of the images: `load_data.decode_u8` (affgw_u8_to_image) against the oracle's
restatement of load_data.py:152-166 and against canvases the reference loader itself produced - bit for bit."""
import numpy as np
import pytest
import torch

from affganwriting_b200 import load_data as LD
from oracle import affgw_oracle as O

pytestmark = pytest.mark.gpu


def test_decode_matches_reference_loader_fixture(golden):
    g = golden("u8_wire.npz")
    for i in range(int(g["count"])):
        wire = O.pad_resized_u8(g[f"u8.{i}"])
        out = LD.decode_u8(torch.from_numpy(wire).cuda())
        assert out.dtype == torch.float32 and tuple(out.shape) == (64, 216)
        assert np.array_equal(out.cpu().numpy(), g[f"ref.{i}"]), i


@pytest.mark.parametrize("shape", [(256,), (1,), (17,), (3, 50, 64, 216), (5, 1, 64, 216), (1, 15), (4099,)])
def test_decode_every_level_and_ragged_sizes(shape):
    rng = np.random.RandomState(11)
    n = int(np.prod(shape))
    u8 = np.arange(n, dtype=np.int64) % 256 if n >= 256 else rng.randint(0, 256, n)
    u8 = rng.permutation(u8).astype(np.uint8).reshape(shape)
    out = LD.decode_u8(torch.from_numpy(u8).cuda())
    assert tuple(out.shape) == shape and np.array_equal(out.cpu().numpy(), O.decode_u8(u8))


def test_decode_rejects_bad_arguments():
    with pytest.raises(RuntimeError):
        LD.decode_u8(torch.zeros(16, dtype=torch.uint8))                         # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        LD.decode_u8(torch.zeros(16, dtype=torch.float32, device="cuda"))
    with pytest.raises(RuntimeError):
        LD.decode_u8(torch.zeros(16, dtype=torch.uint8, device="cuda"), out=torch.zeros(8, device="cuda"))
    assert LD.decode_u8(torch.zeros(0, dtype=torch.uint8, device="cuda")).numel() == 0


def test_batch_to_device_decodes_uint8_images():
    b = O.synthetic_batch(2, 15)
    u8 = torch.randint(0, 256, (2, 15, 64, 216), dtype=torch.uint8)
    host = (b["tr_img"], u8, b["label_xt"], "src")
    dev = LD.batch_to_device(host, torch.device("cuda"))
    assert dev[0].is_cuda and torch.equal(dev[0].cpu(), b["tr_img"])
    assert dev[1].dtype == torch.float32 and np.array_equal(dev[1].cpu().numpy(), O.decode_u8(u8.numpy()))
    assert dev[2].dtype == torch.int64 and dev[3] == "src"


def test_prefetcher_hands_out_batches_in_order():
    dev = torch.device("cuda")
    pf = LD.DevicePrefetcher(dev)
    rng = np.random.RandomState(5)
    hosts = []
    for k in range(5):
        u8 = torch.from_numpy(rng.randint(0, 256, (2, 3, 64, 216)).astype(np.uint8)).pin_memory()
        lab = torch.full((2, 12), k, dtype=torch.int64).pin_memory()
        hosts.append((u8, lab, "tag%d" % k))
    with pytest.raises(RuntimeError):
        pf.get()
    pf.stage(hosts[0])
    for k in range(5):
        img, lab, tag = pf.get()
        if k + 1 < 5:
            pf.stage(hosts[k + 1])
        acc = img.sum() + lab.sum()                 # work on the current stream that reads the slot
        torch.cuda.synchronize()
        assert tag == "tag%d" % k and int(lab[0, 0]) == k
        assert np.array_equal(img.cpu().numpy(), O.decode_u8(hosts[k][0].numpy())) and torch.isfinite(acc)
        pf.release()
    pf.stage(hosts[0])
    pf.stage(hosts[1])
    with pytest.raises(RuntimeError):
        pf.stage(hosts[2])                          # both slots staged
    with pytest.raises(RuntimeError):
        pf2 = LD.DevicePrefetcher(dev)
        pf2.stage(hosts[0]); pf2.get(); pf2.release()
        pf2.stage(hosts[1]); pf2.get(); pf2.release()
        pf2.stage((hosts[0][0][:1], hosts[0][1], "x"))   # slot 0 was sized for another layout
