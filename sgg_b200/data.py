"""Host-side input pipeline with the reference's on-disk formats and batching semantics (train.py:95-226; SURVEY 8 row
f4).  Pure host code (json / PIL / torch CPU ops): it feeds ``SceneGraphGAN.train_from_images``.

Formats (dataset_creation/map_files_to_triples.py):
  * ``vocab.json``            {word: id}                                   (createVocab, :58-117)
  * ``ims_to_triples.json``   {image path: [[subject id, predicate id, object id], ...]}   (mapFromImagesToTriples, :163-175)
  * ``word_embeddings.npy``   [len(vocab), d] float64, U(-0.1, 0.1) where word2vec has no entry   (loadWordEmbeddings, :14-55)
  * ``image_means.txt`` / ``image_stds.txt``   three lines each: per-channel (R, G, B) statistics on the 0..255 scale

Semantics restated from train.py:
  * ``_gatherFiles`` (:114-163): the first 90 % of the images (file order of the json) give (file, triple) pairs, shuffled,
    88 % train / 12 % validation; the last 10 % are test images, every triple of an image one element, the last triple
    repeated until the image has TEST_BATCH_SIZE * TEST_BATCH_MULTIPLIER elements (images without triples are skipped);
  * ``_parseFunction`` (:165-171): decode_jpeg(channels=3) -> resize_images([221, 221]) -> (x - means) / stds, labels one-hot
    (here: kept as ids, the hot path's label format);
  * ``_createSingleDataset`` (:173-190): repeat() -> shuffle(buffer = 10 * batch) -> batch -> every batch repeated
    CRITIC_ITERS + 1 times (one iteration of the trainer consumes one batch for all its steps, which is the same thing).

``tf.image.resize_images`` in TF 1.x is ResizeBilinear with align_corners = False and no half-pixel centres:
source coordinate = destination index * (in / out); ``resize_bilinear_tf1`` restates exactly that (it is NOT
torch.nn.functional.interpolate's convention).
"""
from __future__ import annotations

import json
import queue
import random
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

IMAGE_SIZE = 221               # train.py:168


def load_image_stats(path_to_image_means: str, path_to_image_stds: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """train.py:95-108: one float per line, broadcast over a [221, 221, 3] image."""
    def read(path):
        with open(path) as f:
            vals = [float(line.strip()) for line in f if line.strip()]
        if len(vals) != 3:
            raise ValueError(f"{path}: expected 3 per-channel values, got {len(vals)}")
        return torch.tensor(vals, dtype=torch.float32)
    return read(path_to_image_means), read(path_to_image_stds)


def gather_files(ims_to_triples: Dict[str, List[List[int]]], test_batch_size: int, test_batch_multiplier: int,
                 seed: Optional[int] = None):
    """train.py:114-163.  Returns train_files, train_labels [n,3], val_files, val_labels, test_files, test_labels."""
    keys = list(ims_to_triples.keys())
    cut = int(0.9 * len(keys))
    all_files, all_labels = [], []
    for k in keys[:cut]:
        for t in ims_to_triples[k]:
            all_files.append(k)
            all_labels.append(list(t))
    order = list(range(len(all_files)))
    random.Random(seed).shuffle(order)                                   # sklearn.utils.shuffle(all_files, all_labels)
    all_files = [all_files[i] for i in order]
    all_labels = [all_labels[i] for i in order]
    thr = int(0.88 * len(all_files))
    as_labels = lambda rows: np.asarray(rows, dtype=np.int64).reshape(-1, 3)
    test_files, test_labels = [], []
    need = test_batch_size * test_batch_multiplier
    for k in keys[cut:]:
        triples = ims_to_triples[k]
        if len(triples) == 0:
            continue
        for t in triples:
            test_files.append(k)
            test_labels.append(list(t))
        for _ in range(max(0, need - len(triples))):
            test_files.append(k)
            test_labels.append(list(triples[-1]))
    return (all_files[:thr], as_labels(all_labels[:thr]), all_files[thr:], as_labels(all_labels[thr:]),
            test_files, as_labels(test_labels))


def resize_bilinear_tf1(img: torch.Tensor, out_h: int = IMAGE_SIZE, out_w: int = IMAGE_SIZE) -> torch.Tensor:
    """[H, W, C] float -> [out_h, out_w, C]: TF 1.x ResizeBilinear (align_corners=False, half_pixel_centers=False)."""
    H, W, _ = img.shape

    def axis(n_in, n_out):
        src = torch.arange(n_out, dtype=torch.float32) * (n_in / n_out)
        lo = src.floor().long().clamp_(max=n_in - 1)
        hi = (lo + 1).clamp_(max=n_in - 1)
        return lo, hi, (src - lo.float())
    y0, y1, wy = axis(H, out_h)
    x0, x1, wx = axis(W, out_w)
    top = img[y0][:, x0] + (img[y0][:, x1] - img[y0][:, x0]) * wx[None, :, None]
    bot = img[y1][:, x0] + (img[y1][:, x1] - img[y1][:, x0]) * wx[None, :, None]
    return top + (bot - top) * wy[:, None, None]


def parse_image(path: str, means: torch.Tensor, stds: torch.Tensor) -> torch.Tensor:
    """train.py:165-170: JPEG file -> standardised [221, 221, 3] float32."""
    from PIL import Image
    with Image.open(path) as im:
        rgb = torch.from_numpy(np.asarray(im.convert("RGB"), dtype=np.uint8).copy()).float()
    return (resize_bilinear_tf1(rgb) - means) / stds


def shuffle_buffer(n_items: int, buffer_size: int, rng: random.Random, repeat: bool) -> Iterator[int]:
    """Index stream of ``Dataset.range(n).repeat().shuffle(buffer_size)``: a buffer is filled from the (repeated) input,
    every output is a uniformly drawn buffer slot which is then refilled with the next input element."""
    def source():
        while True:
            yield from range(n_items)
            if not repeat:
                return
    src = source()
    buf: List[int] = []
    for i in src:
        buf.append(i)
        if len(buf) >= buffer_size:
            break
    while buf:
        j = rng.randrange(len(buf))
        out = buf[j]
        nxt = next(src, None)
        if nxt is None:
            buf[j] = buf[-1]
            buf.pop()
        else:
            buf[j] = nxt
        yield out


class ImageBatches:
    """(images [B,221,221,3] float32, labels [B,3] int64) batches as ``_createSingleDataset`` produces them.

    train / val: repeat + shuffle buffer of 10 batches (train.py:175-177); test: file order, one pass (train.py:299).
    rank / world: rank r takes batches r, r + world, ... of the stream (the reference is single-process).
    A short final batch of the one-pass mode is dropped (the engine's batch size is fixed).
    prefetch: batches decoded ahead by a background thread (``d.prefetch(...)``, train.py:188), so that JPEG decoding
    overlaps the consumer's GPU work; 0 decodes in the consumer's thread."""

    def __init__(self, files: Sequence[str], labels: np.ndarray, batch_size: int, means: torch.Tensor, stds: torch.Tensor,
                 shuffle: bool = True, repeat: bool = True, seed: int = 0, rank: int = 0, world: int = 1, workers: int = 8,
                 pin_memory: bool = False, prefetch: int = 2):
        if len(files) != len(labels):
            raise ValueError("files / labels length mismatch")
        self.files, self.labels = list(files), np.asarray(labels, dtype=np.int64).reshape(-1, 3)
        self.B, self.means, self.stds = int(batch_size), means, stds
        self.shuffle, self.repeat, self.seed = shuffle, repeat, seed
        self.rank, self.world, self.workers, self.pin = rank, world, max(1, workers), pin_memory
        self.prefetch = int(prefetch)

    def _index_stream(self) -> Iterator[int]:
        if self.shuffle:
            return shuffle_buffer(len(self.files), self.B * 10, random.Random(self.seed), self.repeat)
        def plain():
            while True:
                yield from range(len(self.files))
                if not self.repeat:
                    return
        return plain()

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        if self.prefetch <= 0:
            yield from self._batches()
            return
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop, end = threading.Event(), object()

        def put(item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def produce():
            try:
                for item in self._batches():
                    if not put(item):
                        return
                put(end)
            except BaseException as exc:          # decoding errors surface in the consumer
                put(exc)

        threading.Thread(target=produce, daemon=True).start()
        try:
            while True:
                item = q.get()
                if item is end:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()                            # the consumer stopped early: release the producer

    def _batches(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        if len(self.files) == 0:
            return
        stream = self._index_stream()
        n = 0
        with ThreadPoolExecutor(self.workers) as pool:
            while True:
                idx = []
                for i in stream:
                    idx.append(i)
                    if len(idx) == self.B:
                        break
                if len(idx) < self.B:
                    return
                mine = (n % self.world) == self.rank
                n += 1
                if not mine:
                    continue
                images = torch.stack(list(pool.map(lambda i: parse_image(self.files[i], self.means, self.stds), idx)))
                labels = torch.from_numpy(self.labels[idx])
                if self.pin and torch.cuda.is_available():
                    images, labels = images.pin_memory(), labels.pin_memory()
                yield images, labels


def load_dataset(path_to_ims_to_triples: str, path_to_image_means: str, path_to_image_stds: str, batch_size: int,
                 test_batch_size: Optional[int] = None, test_batch_multiplier: int = 8, seed: int = 0, rank: int = 0,
                 world: int = 1, eval_batch_size: Optional[int] = None):
    """train.py:192-212 ``_loadDatasets``: the three batch streams plus the iteration bookkeeping of ``_gatherFiles``.
    test_batch_size sizes the per-image padding of the test list (train.py:30: batch_size / 2); eval_batch_size is the
    batch the validation / test streams are cut into (the reference uses batch_size / 2 as well; an engine with a fixed
    batch passes its own)."""
    with open(path_to_ims_to_triples) as f:
        ims_to_triples = json.load(f)
    if test_batch_size is None:
        test_batch_size = max(1, batch_size // 2)
    if eval_batch_size is None:
        eval_batch_size = test_batch_size
    means, stds = load_image_stats(path_to_image_means, path_to_image_stds)
    trf, trl, vaf, val, tef, tel = gather_files(ims_to_triples, test_batch_size, test_batch_multiplier, seed)
    if not len(trf) > len(vaf):
        raise ValueError("the training split must be larger than the validation split (train.py:161)")
    return {
        "train": ImageBatches(trf, trl, batch_size, means, stds, True, True, seed, rank, world),
        "val": ImageBatches(vaf, val, eval_batch_size, means, stds, True, True, seed + 1, rank, world),
        "test": ImageBatches(tef, tel, eval_batch_size, means, stds, False, False, seed, rank, world),
        "max_iterations": 5 * len(trf),                 # train.py:153
        "write_iterations": 10,                         # train.py:155
        "validate_iterations": int(len(trf) / 50),      # train.py:156
    }
