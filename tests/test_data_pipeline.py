"""Host-side input pipeline (sgg_b200/data.py) against literal restatements of train.py:114-190 written
here as loops: the split rules of _gatherFiles, TF 1.x bilinear resizing, standardisation, the shuffle-buffer stream and
the per-rank batch split.  CPU only."""
import json
import random

import numpy as np
import pytest
import torch

from sgg_b200 import data as D


def _dataset(n_images, rng):
    return {f"/images/{i}.jpg": [[rng.randrange(50), rng.randrange(50), rng.randrange(50)] for _ in range(rng.randrange(0, 7))]
            for i in range(n_images)}


def test_gather_files_follows_the_reference_split_rules():
    rng = random.Random(3)
    ims = _dataset(40, rng)
    trf, trl, vaf, val, tef, tel = D.gather_files(ims, test_batch_size=4, test_batch_multiplier=8, seed=1)
    keys = list(ims)
    head, tail = keys[:36], keys[36:]                                   # int(0.9 * 40)
    pairs = [(k, tuple(t)) for k in head for t in ims[k]]
    assert len(trf) == int(0.88 * len(pairs)) and len(trf) + len(vaf) == len(pairs)             # train.py:139-143
    got = sorted(list(zip(trf, map(tuple, trl.tolist()))) + list(zip(vaf, map(tuple, val.tolist()))))
    assert got == sorted(pairs)                                          # a permutation of every (file, triple) pair
    assert trl.shape[1] == 3 and trl.dtype == np.int64
    assert not set(tef) & set(head)                                      # test images come from the last 10 % only
    # train.py:147-158: every triple once, then the last triple until the image has TEST_BATCH_SIZE * MULTIPLIER elements
    pos = 0
    for k in tail:
        triples = ims[k]
        if not triples:
            assert k not in tef
            continue
        n = max(len(triples), 32)
        assert tef[pos:pos + n] == [k] * n
        assert tel[pos:pos + len(triples)].tolist() == triples
        assert all(row == triples[-1] for row in tel[pos + len(triples):pos + n].tolist())
        pos += n
    assert pos == len(tef)


def _resize_loops(img, oh, ow):
    """ResizeBilinear of TF 1.x, align_corners=False, no half-pixel centres, as scalar loops."""
    H, W, C = img.shape
    out = np.zeros((oh, ow, C), dtype=np.float64)
    for y in range(oh):
        fy = y * (H / oh)
        y0 = min(int(np.floor(fy)), H - 1); y1 = min(y0 + 1, H - 1); wy = fy - y0
        for x in range(ow):
            fx = x * (W / ow)
            x0 = min(int(np.floor(fx)), W - 1); x1 = min(x0 + 1, W - 1); wx = fx - x0
            top = img[y0, x0] + (img[y0, x1] - img[y0, x0]) * wx
            bot = img[y1, x0] + (img[y1, x1] - img[y1, x0]) * wx
            out[y, x] = top + (bot - top) * wy
    return out


@pytest.mark.parametrize("shape,out", [((7, 5, 3), (11, 9)), ((30, 40, 3), (13, 17)), ((4, 4, 3), (4, 4))])
def test_resize_matches_the_tf1_bilinear_rule(shape, out):
    img = torch.rand(*shape, generator=torch.Generator().manual_seed(0)) * 255
    got = D.resize_bilinear_tf1(img, *out)
    ref = _resize_loops(img.double().numpy(), *out)
    assert got.shape == (out[0], out[1], 3)
    assert np.abs(got.numpy() - ref).max() < 1e-3          # fp32 on 0..255 values
    if shape[:2] == out:
        assert torch.equal(got, img)                        # identity when the size is unchanged
    assert torch.equal(got[0, 0], img[0, 0])                # destination (0, 0) samples source (0, 0): no half-pixel shift


def test_parse_image_decodes_resizes_and_standardises(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(0)
    arr = rng.randint(0, 256, size=(60, 80, 3), dtype=np.uint8)
    path = str(tmp_path / "a.jpg")
    Image.fromarray(arr).save(path, quality=95)
    (tmp_path / "means.txt").write_text("119.619349848\n115.116956701\n106.136688569\n")    # dataset_creation/image_means.txt
    (tmp_path / "stds.txt").write_text("30.3701687507\n30.491321632\n36.7036557611\n")
    means, stds = D.load_image_stats(str(tmp_path / "means.txt"), str(tmp_path / "stds.txt"))
    got = D.parse_image(path, means, stds)
    assert got.shape == (221, 221, 3) and got.dtype == torch.float32
    with Image.open(path) as im:
        decoded = np.asarray(im.convert("RGB")).astype(np.float64)
    ref = (_resize_loops(decoded, 221, 221) - means.double().numpy()) / stds.double().numpy()
    assert np.abs(got.numpy() - ref).max() < 1e-4
    # grey-scale files come out with three channels (decode_jpeg(channels=3))
    Image.fromarray(arr[:, :, 0]).save(str(tmp_path / "g.jpg"))
    assert D.parse_image(str(tmp_path / "g.jpg"), means, stds).shape == (221, 221, 3)
    with pytest.raises(ValueError):
        (tmp_path / "bad.txt").write_text("1\n2\n")
        D.load_image_stats(str(tmp_path / "bad.txt"), str(tmp_path / "stds.txt"))


def test_shuffle_buffer_is_a_windowed_permutation():
    n, buf = 100, 16
    once = list(D.shuffle_buffer(n, buf, random.Random(0), repeat=False))
    assert sorted(once) == list(range(n)) and once != list(range(n))
    # element i enters the buffer when output (i - buf + 1) is drawn, so it cannot appear earlier than that
    for pos, i in enumerate(once):
        assert pos >= i - buf + 1
    stream = D.shuffle_buffer(n, buf, random.Random(1), repeat=True)
    many = [next(stream) for _ in range(5 * n)]
    counts = np.bincount(many, minlength=n)
    assert counts.min() >= 3 and counts.max() <= 7          # repeat(): every element keeps coming back
    assert list(D.shuffle_buffer(5, 64, random.Random(0), repeat=False)).__len__() == 5      # buffer larger than the data


def test_batches_have_the_engine_format_and_split_over_ranks(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(1)
    files, labels = [], []
    for i in range(12):
        p = str(tmp_path / f"{i}.jpg")
        Image.fromarray(rng.randint(0, 256, size=(20 + i, 30, 3), dtype=np.uint8)).save(p)
        files.append(p); labels.append([i, i + 1, i + 2])
    means, stds = torch.tensor([100.0, 110.0, 120.0]), torch.tensor([30.0, 31.0, 32.0])
    one_pass = list(D.ImageBatches(files, np.array(labels), 4, means, stds, shuffle=False, repeat=False, workers=2))
    assert len(one_pass) == 3
    for b, (im, lb) in enumerate(one_pass):
        assert im.shape == (4, 221, 221, 3) and im.dtype == torch.float32 and lb.dtype == torch.int64
        assert lb.tolist() == labels[4 * b:4 * b + 4]
        assert torch.equal(im[1], D.parse_image(files[4 * b + 1], means, stds))
    # a short last batch is dropped; ranks take alternating batches of the same stream
    assert len(list(D.ImageBatches(files[:10], np.array(labels[:10]), 4, means, stds, shuffle=False, repeat=False))) == 2
    r0 = list(D.ImageBatches(files, np.array(labels), 2, means, stds, shuffle=True, repeat=False, seed=5, rank=0, world=2))
    r1 = list(D.ImageBatches(files, np.array(labels), 2, means, stds, shuffle=True, repeat=False, seed=5, rank=1, world=2))
    seen = [tuple(row) for _, lb in r0 + r1 for row in lb.tolist()]
    assert len(r0) == 3 and len(r1) == 3 and sorted(seen) == sorted(map(tuple, labels))
    assert list(D.ImageBatches([], np.zeros((0, 3)), 2, means, stds)) == []


def test_load_dataset_wires_the_three_streams(tmp_path):
    from PIL import Image
    rng = random.Random(2)
    ims = {}
    for i in range(20):
        p = str(tmp_path / f"{i}.jpg")
        Image.fromarray(np.full((16, 16, 3), 10 * i % 255, dtype=np.uint8)).save(p)
        ims[p] = [[rng.randrange(9), rng.randrange(9), rng.randrange(9)] for _ in range(1 + i % 3)]
    (tmp_path / "ims.json").write_text(json.dumps(ims))
    (tmp_path / "m.txt").write_text("1\n2\n3\n"); (tmp_path / "s.txt").write_text("4\n5\n6\n")
    ds = D.load_dataset(str(tmp_path / "ims.json"), str(tmp_path / "m.txt"), str(tmp_path / "s.txt"), batch_size=4)
    n_pairs = sum(len(ims[k]) for k in list(ims)[:18])
    assert ds["max_iterations"] == 5 * int(0.88 * n_pairs) and ds["write_iterations"] == 10      # train.py:153-156
    assert ds["validate_iterations"] == int(int(0.88 * n_pairs) / 50)
    im, lb = next(iter(ds["train"]))
    assert im.shape == (4, 221, 221, 3) and lb.shape == (4, 3)
    test_batches = list(ds["test"])                        # 2 test images x (batch_size / 2 = 2) x 8 = 16 elements each
    assert len(test_batches) == 16 and all(lb.shape == (2, 3) for _, lb in test_batches)
    first_image_rows = [tuple(r) for _, lb in test_batches[:8] for r in lb.tolist()]
    assert set(first_image_rows) == set(map(tuple, ims[list(ims)[18]]))


def test_prefetch_thread_keeps_order_stops_early_and_reports_errors(tmp_path):
    import threading
    import time

    from PIL import Image
    files, labels = [], []
    for i in range(8):
        p = str(tmp_path / f"{i}.jpg")
        Image.fromarray(np.full((12, 12, 3), 20 * i, dtype=np.uint8)).save(p)
        files.append(p); labels.append([i, i, i])
    means, stds = torch.zeros(3), torch.ones(3)
    direct = list(D.ImageBatches(files, np.array(labels), 2, means, stds, shuffle=True, repeat=False, seed=3, prefetch=0))
    ahead = list(D.ImageBatches(files, np.array(labels), 2, means, stds, shuffle=True, repeat=False, seed=3, prefetch=3))
    assert len(direct) == len(ahead) == 4
    for (a, la), (b, lb) in zip(direct, ahead):
        assert torch.equal(a, b) and torch.equal(la, lb)
    # an endless stream abandoned after two batches: the producer thread goes away
    before = threading.active_count()
    it = iter(D.ImageBatches(files, np.array(labels), 2, means, stds, shuffle=True, repeat=True, prefetch=2))
    next(it); next(it)
    it.close()
    deadline = time.time() + 5
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before
    # a missing file raises in the consumer, not silently in the background
    bad = D.ImageBatches(files[:2] + [str(tmp_path / "missing.jpg"), files[3]], np.array(labels[:4]), 2, means, stds,
                         shuffle=False, repeat=False, prefetch=2)
    with pytest.raises(FileNotFoundError):
        list(bad)
