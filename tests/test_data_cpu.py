"""CPU test of the batched-file reader: same file list and channel split as the reference's
WaymoDataset.get_batch on a synthetic on-disk dataset, ring slots stay valid while read ahead."""
import os

import pytest
import torch

from dmmfods_b200 import data
from oracle import ref_shim


def _make_dataset(root, n=5, B=2, H=8, W=12):
    torch.manual_seed(0)
    d = os.path.join(root, "train", "subset0")
    os.makedirs(os.path.join(d, "labels"))
    batches = []
    for i in range(n):
        t = torch.randn(B, 7, H, W)
        torch.save(t, os.path.join(d, str(i)))
        batches.append(t)
    return batches


def test_file_list_and_split_match_the_reference(tmp_path):
    root = str(tmp_path)
    batches = _make_dataset(root)
    files = data.list_batch_files(root, "train")
    assert sorted(files) == sorted(os.path.join("train", "subset0", str(i)) for i in range(5))
    ring = data.BatchFileRing(root, files, depth=2, pin=False)
    got = []
    for img, lid, hm in ring:
        got.append((img.clone(), lid.clone(), hm.clone()))
    assert len(got) == len(ring) == 5
    for rel, (img, lid, hm) in zip(files, got):
        t = batches[int(os.path.basename(rel))]
        assert torch.equal(img, t[:, :3]) and torch.equal(lid, t[:, 3:4]) and torch.equal(hm, t[:, 4:])
        assert img.shape == (2, 3, 8, 12) and lid.shape == (2, 1, 8, 12) and hm.shape == (2, 3, 8, 12)


def test_ring_slots_stay_valid_while_reading_ahead(tmp_path):
    root = str(tmp_path)
    batches = _make_dataset(root, n=6)
    files = sorted(data.list_batch_files(root, "train"), key=lambda f: int(os.path.basename(f)))
    it = iter(data.BatchFileRing(root, files, depth=3, pin=False, epochs=2))
    a = next(it)
    b = next(it)        # the reader is now up to depth batches ahead; `a` and `b` must still hold batches 0 and 1
    assert torch.equal(a[0], batches[0][:, :3]) and torch.equal(b[0], batches[1][:, :3])
    rest = list(it)
    assert len(rest) == 10 and torch.equal(rest[-1][2], batches[5][:, 4:])


def test_bad_batch_shape_raises(tmp_path):
    root = str(tmp_path)
    d = os.path.join(root, "train", "s")
    os.makedirs(d)
    torch.save(torch.zeros(2, 6, 4, 4), os.path.join(d, "0"))
    with pytest.raises(ValueError):
        list(data.BatchFileRing(root, data.list_batch_files(root, "train"), pin=False))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_split_matches_live_reference_get_batch(tmp_path):
    import importlib
    ref_shim.install()
    wd = importlib.import_module("dmmfods.datasets.WaymoData")
    root = str(tmp_path)
    batches = _make_dataset(root, n=2)
    ds = wd.WaymoDataset.__new__(wd.WaymoDataset)           # get_batch only needs root + files
    ds.root, ds.files = root, data.list_batch_files(root, "train")
    for i, (img, lid, hm) in enumerate(data.BatchFileRing(root, ds.files, pin=False)):
        r_img, r_lid, r_hm = ds.get_batch(i)
        assert torch.equal(img, r_img) and torch.equal(lid, r_lid) and torch.equal(hm, r_hm)
