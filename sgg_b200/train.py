"""``SceneGraphGAN`` trainer with the reference's interface (train.py:17-422) over the B200 hot path.

Kept from the reference: the constructor signature (train.py:23-24), ``_Generator`` / ``_Discriminator`` /
``train`` method names, the CLI flag names (train.py:399-412), the step schedule (CRITIC_ITERS D steps then one G
step on the same batch, train.py:185-187,362-368), the losses (train.py:245-253) and the two Adam optimisers
(train.py:258-259).  Not reproduced (outside the hot path): the tf.data JPEG pipeline, summaries.  The conv
front-end (gen:29-68) exists as a caller-side module on library convolutions (``frontend.py``, ``train_from_images``);
the hot path proper starts at the annotation grid.  Documented deviations: the R@k ranking bug of train.py:320 (see ``test``); the reference passes batch_size / critic_iters to
the constructor in swapped order (train.py:420-421 vs 23-24) -- here each flag reaches its own parameter; the
constructor does not delete the contents of the checkpoint / summary directories (train.py:50-53).
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Iterable, Optional

import torch

from .architectures import Discriminator, Generator
from .trainer import HotPathTrainer


class SceneGraphGAN(object):
    def __init__(self, checkpoints_dir, summaries_dir, path_to_ims_to_triples, path_to_vocab, path_to_word_embeddings,
                 path_to_image_means, path_to_image_stds, critic_iters, batch_size, lambda_, resume,
                 vocab_size: Optional[int] = None, n_triples: int = 1, seed: int = 0, allow_synthetic: bool = True):
        # Hyperparameters (train.py:27-32)
        self.CRITIC_ITERS = int(critic_iters)
        self.BATCH_SIZE = int(batch_size)
        self.LAMBDA = float(lambda_)
        self.resume = bool(resume)
        self.allow_synthetic = bool(allow_synthetic)
        self.checkpoints_dir, self.summaries_dir = checkpoints_dir, summaries_dir
        self.path_to_ims_to_triples = path_to_ims_to_triples
        self.path_to_image_means, self.path_to_image_stds = path_to_image_means, path_to_image_stds
        for d in (checkpoints_dir, summaries_dir):
            if d and not os.path.exists(d):
                os.makedirs(d)
        # vocabulary and embeddings (train.py:55-63); both optional for synthetic runs
        self.vocab = None
        if path_to_vocab and os.path.exists(path_to_vocab):
            with open(path_to_vocab) as f:
                self.vocab = json.load(f)
            vocab_size = len(self.vocab)
            self.reverse_vocab = {y: x for x, y in self.vocab.items()}      # train.py:77
        if vocab_size is None:
            raise ValueError("need path_to_vocab or vocab_size")
        self.embeddings = None
        if path_to_word_embeddings and os.path.exists(path_to_word_embeddings):
            import numpy as np
            self.embeddings = torch.from_numpy(np.load(path_to_word_embeddings)).float()
            if self.embeddings.shape[0] != vocab_size:
                raise ValueError("word_embeddings rows != vocabulary size")
        self.ims_to_triples = None
        if path_to_ims_to_triples and os.path.exists(path_to_ims_to_triples):
            with open(path_to_ims_to_triples) as f:
                self.ims_to_triples = json.load(f)
        self.n_steps = 3 * int(n_triples)                                    # gen:85 hard-codes 3 = one triple
        E = self.embeddings.shape[1] if self.embeddings is not None else 300
        self.trainer = HotPathTrainer(self.BATCH_SIZE, self.n_steps, vocab_size, critic_iters=self.CRITIC_ITERS,
                                      lam=self.LAMBDA, embed_dim=E, seed=seed, embedding=self.embeddings)
        # train.py:65-72: Generator(len(vocab)); Discriminator(len(vocab), W) with W initialised from the embeddings
        self.g = Generator(vocab_size, self.n_steps)
        self.d = Discriminator(vocab_size, self.trainer.eng.d.views()["Discriminator/W"], self.n_steps)
        self.g._attach(self.trainer.eng)
        self.d._attach(self.trainer.eng)
        self.front = None                      # FrontEndTrainer, created by the first train_from_images() / on load
        self._seed = seed
        if self.resume:
            self._loadModel()

    # ------------------------------------------------------------------ train.py:85-93
    def _Generator(self, images, is_training=True):
        return self.g.build_generator(images, is_training)

    def _Discriminator(self, triple_input, images, is_training=True):
        return self.d.build_discriminator(triple_input, images, is_training)

    def _front(self):
        """The conv front-ends (gen:29-68 / disc:29-68) and their optimisers; library convolutions, see frontend.py."""
        if self.front is None:
            from .frontend import FrontEndTrainer
            self.front = FrontEndTrainer(self.trainer, seed=self._seed)
        return self.front

    # ------------------------------------------------------------------ train.py:288-292 (never called there)
    def _ckpt(self):
        return os.path.join(self.checkpoints_dir, "model.ckpt.pt")

    def _saveModel(self, itr=None):
        """Collective under data parallelism: every rank calls it.  With the row-sharded attention projection each
        rank only maintains its own rows of attention_perceptron/kernel, so the rows are all-gathered first; rank 0
        alone writes the file, then all ranks meet at a barrier."""
        tr = self.trainer
        tr.gather_sharded()
        e = tr.eng
        if tr.rank == 0:
            ck = {"generator": e.g.state_dict(), "discriminator": e.d.state_dict(),
                  "adam": {"g": (e.g.m.cpu(), e.g.v.cpu(), e.g.step), "d": (e.d.m.cpu(), e.d.v.cpu(), e.d.step)},
                  "iterations": int(e.counters.item())}
            if self.front is not None:      # conv variables under their TF names (HWIO kernels) + their Adam state
                f = self.front
                ck["front_end"] = {k: v.cpu().clone() for net in (f.fg, f.fd) for k, v in net.tf_variables().items()}
                ck["front_end_adam"] = {"g": f.adam_fg.state(), "d": f.adam_fd.state()}
            torch.save(ck, self._ckpt())
        if tr.dist is not None:
            tr.dist.barrier(group=tr.pg)

    def load_tf_checkpoint(self, prefix: str):
        """Restores a checkpoint written by the reference's ``tf.train.Saver`` (train.py:280,288-292; TF V2 format
        ``<prefix>.index`` + ``<prefix>.data-*``): the hot-path variables go into the parameter buckets under their TF
        names, the ``/Adam`` and ``/Adam_1`` slots into the Adam moments, ``beta1_power`` / ``beta1_power_1`` give the
        optimiser steps (generator's optimiser is created first, train.py:258-259).  Variables of the convolutional
        front-end are ignored (outside the hot path).  Returns the names that were not consumed."""
        from . import tf_checkpoint as T
        parts = T.split_for_buckets(T.read_checkpoint(prefix), beta1=self.trainer.beta1)
        e = self.trainer.eng
        for bucket, key in ((e.g, "generator"), (e.d, "discriminator")):
            bucket.load_state_dict({k: torch.from_numpy(v) for k, v in parts[key].items()})
            mv, vv = bucket._views(bucket.m), bucket._views(bucket.v)
            for k in bucket.views():
                if k in parts["adam_m"]:
                    mv[k].copy_(torch.from_numpy(parts["adam_m"][k]).reshape(mv[k].shape))
                if k in parts["adam_v"]:
                    vv[k].copy_(torch.from_numpy(parts["adam_v"][k]).reshape(vv[k].shape))
        e.g.step = parts["step"].get("beta1_power", 0)
        e.d.step = parts["step"].get("beta1_power_1", 0)
        e.counters.fill_(e.g.step)
        other = parts["other"]
        if any("/conv2d" in k for k in other):     # the conv front-end's variables (gen:29-68), if the caller trains it
            f = self._front()
            for net in (f.fg, f.fd):
                for name in net.load_tf_variables({k: torch.from_numpy(v) for k, v in other.items()}, strict=False):
                    other.pop(name)
        return sorted(other)

    def _loadModel(self):
        if not os.path.exists(self._ckpt()):      # a reference (TensorFlow) checkpoint in the directory?
            tf_prefix = os.path.join(self.checkpoints_dir, "model.ckpt")
            if os.path.exists(tf_prefix + ".index"):
                self.load_tf_checkpoint(tf_prefix)
                return
        ck = torch.load(self._ckpt(), map_location="cpu")
        e = self.trainer.eng
        e.g.load_state_dict(ck["generator"])
        e.d.load_state_dict(ck["discriminator"])
        for b, key in ((e.g, "g"), (e.d, "d")):
            m, v, step = ck["adam"][key]
            b.m.copy_(m); b.v.copy_(v); b.step = step
        e.counters.fill_(ck["iterations"])
        if "front_end" in ck:
            f = self._front()
            f.fg.load_tf_variables(ck["front_end"])
            f.fd.load_tf_variables(ck["front_end"])
            f.adam_fg.load_state(ck["front_end_adam"]["g"])
            f.adam_fd.load_state(ck["front_end_adam"]["d"])

    # ------------------------------------------------------------------ train.py:341-388
    def train(self, batches: Optional[Iterable] = None, max_iterations: Optional[int] = None, log_every: int = 10):
        """``batches`` yields (annotations_for_G, annotations_for_D, labels) host tensors: bf16/float [B,14,14,512]
        twice and int64 [B, n_steps] class ids (the one-hot of train.py:173).  Without it, synthetic batches of
        the configured shape are generated."""
        tr = self.trainer
        if batches is None:
            if self.ims_to_triples is not None and not self.allow_synthetic:
                # the reference's own mode: JPEG files -> conv front-ends -> recurrent half (train.py:341-388)
                have_stats = all(p and os.path.exists(p) for p in (self.path_to_image_means, self.path_to_image_stds))
                if not have_stats or tr.R != 196 or tr.world > 1:
                    raise RuntimeError(
                        "SceneGraphGAN.train(): an ims_to_triples file was loaded but training from its images needs "
                        "image_means.txt / image_stds.txt, 196 regions and a single GPU (the conv front-end is a caller-side "
                        "module on library convolutions). Pass annotation `batches`, or --synthetic.")
                from .data import load_dataset
                ds = load_dataset(self.path_to_ims_to_triples, self.path_to_image_means, self.path_to_image_stds, tr.B,
                                  test_batch_multiplier=self.TEST_BATCH_MULTIPLIER, seed=self._seed, eval_batch_size=tr.B)
                self.dataset = ds
                logs = self.train_from_images(ds["train"], max_iterations or ds["max_iterations"], ds["val"],
                                              ds["validate_iterations"])
                self.last_image_log = logs[-1] if logs else None
                return len(logs)
            batches = synthetic_batches(tr.B, tr.T, tr.V, tr.R, max_iterations or 100, seed=1234 + tr.rank)
        log_path = os.path.join(self.summaries_dir, "train_log.jsonl") if (self.summaries_dir and tr.rank == 0) else None
        t0, n = time.time(), 0
        out = open(log_path, "a") if log_path else None
        try:
            for losses in tr.fit(_to_host_batches(batches, tr)):
                n += 1
                if out and (n % log_every == 0 or n == 1):
                    out.write(json.dumps({"iteration": n, "images_per_s": n * tr.B * tr.world / (time.time() - t0), **losses}) + "\n")
                if max_iterations and n >= max_iterations:
                    break
        finally:
            if out:
                out.close()
        return n

    def train_from_images(self, batches: Iterable, max_iterations: Optional[int] = None, val_batches: Optional[Iterable] = None,
                          validate_every: Optional[int] = None):
        """The reference's loop on pixels (train.py:362-368 with gen:29-68 / disc:29-68 in the graph): ``batches`` yields
        (images [B,221,221,3] float, standardised as train.py:170 does; labels [B, n_steps] int64).  The convolutional
        front-ends are library calls (frontend.py); every step's recurrent half runs through the step-level C ABI
        (sgg_disc_step / sgg_gen_step with the annotation adjoint as an extra output), not through the graph-captured
        sgg_train_iteration -- the conv variables change between the steps.  Single GPU.
        With ``val_batches`` the reference's early stopping applies (train.py:378-388): every ``validate_every`` iterations
        disc_cost of one held-out batch is compared with the previous one; three consecutive increases end the run.
        Returns the per-iteration logs."""
        if self.trainer.world > 1:
            raise RuntimeError("train_from_images: the conv front-end is not data-parallel in this build")
        f = self._front()
        dev = self.trainer.eng.device
        put = lambda im, lb: (im.to(device=dev, dtype=torch.float32), lb.to(device=dev, dtype=torch.int64).contiguous())
        val_it = iter(val_batches) if val_batches is not None else None
        logs, last_loss, worse = [], float("inf"), 0
        for itr, (images, labels) in enumerate(batches):
            logs.append(f.iteration(*put(images, labels)))
            if val_it is not None and validate_every and itr % validate_every == 0:
                loss = f.validation_cost(*put(*next(val_it)))
                logs[-1]["val_disc_cost"] = loss
                worse = worse + 1 if last_loss < loss else 0
                if worse == 3:
                    break
                last_loss = loss
            if max_iterations and len(logs) >= max_iterations:
                break
        return logs

    def test_from_images(self, batches: Iterable, multiplier: Optional[int] = None, out_path: Optional[str] = None):
        """train.py:297-335 on pixels: the annotations of every test batch come from the conv front-ends, then ``test``."""
        f = self._front()
        dev = self.trainer.eng.device

        def annotated():
            for images, labels in batches:
                ag, ad = f.annotations(images.to(device=dev, dtype=torch.float32))
                yield ag, ad, labels
        return self.test(annotated(), multiplier, out_path)

    @staticmethod
    def _recall(fake, real, N):
        """train.py:294-295: |set(fake triples) & set(real triples)| / N."""
        return float(len(set(map(tuple, fake)).intersection(set(map(tuple, real))))) / N

    TEST_BATCH_MULTIPLIER = 8                                                  # train.py:31

    def test(self, batches: Iterable, multiplier: Optional[int] = None, out_path: Optional[str] = None):
        """R@50 / R@100 evaluation of train.py:297-335 on the hot path: for every test batch, ``multiplier`` generator
        passes with fresh noise (train.py:311: TEST_BATCH_MULTIPLIER) give fake triples ``argmax(logits)`` (train.py:270)
        and discriminator scores ``mean_t D(logits, images)`` (train.py:272,314); the fakes are ranked by score and the
        top 50 / 100 are intersected with the batch's real triples (train.py:326-327).  ``batches`` yields
        (annotations_for_G, annotations_for_D, labels [B, n_steps]).
        Documented deviation: the reference ranks with ``score_accumulator.argsort()`` on an [N,1] array, which sorts
        each 1-element row and returns all zeros; here the fakes are sorted by descending critic score.
        Returns (mean R@50, mean R@100) and, like the reference, writes them to ``recalls.txt``."""
        if multiplier is None:
            multiplier = self.TEST_BATCH_MULTIPLIER
        self.trainer.gather_sharded()      # evaluation reads the full attention kernel (collective when world > 1)
        e = self.trainer.eng
        r50, r100 = [], []
        for ag, ad, lb in batches:
            dev = e.device
            ag = ag.to(device=dev, dtype=torch.bfloat16).contiguous()
            ad = ad.to(device=dev, dtype=torch.bfloat16).contiguous()
            lb = lb.to(device=dev, dtype=torch.int64).contiguous()
            e.set_batch(ag.view(e.B, e.R, 512), ad.view(e.B, e.R, 512), lb)
            fakes, scores = [], []
            for _ in range(multiplier):
                e.sample_noise()
                logits = e.gen_forward()                                   # self.fake_inputs (train.py:269)
                fakes.append(logits.argmax(dim=-1))                        # self.fake_triples (train.py:270)
                scores.append(e.disc_forward(logits).mean(dim=1))          # np.mean(disc_scores, axis=1) (train.py:314)
            fake = torch.cat(fakes).cpu().numpy()
            score = torch.cat(scores).cpu().numpy()
            real = lb.cpu().numpy()
            order = (-score).argsort(kind="stable")
            r50.append(self._recall(fake[order[:50]], real, 50.0))
            r100.append(self._recall(fake[order[:100]], real, 100.0))
        m50 = float(sum(r50) / len(r50)) if r50 else 0.0
        m100 = float(sum(r100) / len(r100)) if r100 else 0.0
        path = out_path or "recalls.txt"
        if self.trainer.rank == 0:
            with open(path, "w") as f:                                     # train.py:333-335
                f.write("{}\n{}".format(m50, m100))
        return m50, m100


def synthetic_batches(B, T, V, R, n, seed=1234):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n):
        yield (torch.randn(B, R, 512, generator=g).to(torch.bfloat16), torch.randn(B, R, 512, generator=g).to(torch.bfloat16),
               torch.randint(0, V, (B, T), generator=g))


def _to_host_batches(batches, tr):
    for ag, ad, lb in batches:
        yield (ag.to(torch.bfloat16).contiguous().pin_memory(), ad.to(torch.bfloat16).contiguous().pin_memory(),
               lb.to(torch.int64).contiguous().pin_memory())


def main(argv=None):
    p = argparse.ArgumentParser()   # flag names of train.py:399-412
    p.add_argument("--checkpoints_dir", default="./checkpoints")
    p.add_argument("--summaries_dir", default="./logs")
    p.add_argument("--path_to_ims_to_triples", default="./dataset_creation/ims_to_triples.json")
    p.add_argument("--path_to_vocab", default="./dataset_creation/vocab.json")
    p.add_argument("--path_to_word_embeddings", default="./dataset_creation/word_embeddings.npy")
    p.add_argument("--path_to_image_means", default="./dataset_creation/image_means.txt")
    p.add_argument("--path_to_image_stds", default="./dataset_creation/image_stds.txt")
    p.add_argument("--batch_size", default=64, type=int)
    p.add_argument("--critic_iters", default=10, type=int)
    p.add_argument("--lambda", dest="lambda_", default=10, type=int)
    p.add_argument("--resume", default=False, type=bool)
    p.add_argument("--GPU", default="0")
    # additions for the hot-path build
    p.add_argument("--vocab_size", type=int, default=None, help="synthetic vocabulary size when no vocab.json exists")
    p.add_argument("--n_triples", type=int, default=1)
    p.add_argument("--iterations", type=int, default=100)
    p.add_argument("--synthetic", action="store_true",
                   help="train on synthetic annotation batches even when dataset files are present")
    a = p.parse_args(argv)
    distributed = "LOCAL_RANK" in os.environ
    if distributed:   # torchrun: one process per GPU over NCCL (the reference is single-GPU, train.py:417-418)
        import torch.distributed as dist
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", a.GPU)          # train.py:417-418
    gan = SceneGraphGAN(a.checkpoints_dir, a.summaries_dir, a.path_to_ims_to_triples, a.path_to_vocab,
                        a.path_to_word_embeddings, a.path_to_image_means, a.path_to_image_stds,
                        critic_iters=a.critic_iters, batch_size=a.batch_size, lambda_=a.lambda_, resume=a.resume,
                        vocab_size=a.vocab_size, n_triples=a.n_triples, allow_synthetic=a.synthetic)
    if gan.ims_to_triples is None and not a.synthetic:
        print("train.py: no dataset files found; training on synthetic annotation batches", flush=True)
        gan.allow_synthetic = True
    n = gan.train(max_iterations=a.iterations)
    gan._saveModel()                                                   # collective; rank 0 writes
    image_log = getattr(gan, "last_image_log", None)      # set when the run went through the conv front-ends
    if image_log is not None:
        print(json.dumps({"iterations": n, "disc_cost": image_log["disc_cost"][-1] if image_log["disc_cost"] else None,
                          "gen_cost": image_log["gen_cost"]}))
    elif gan.trainer.rank == 0:
        print(json.dumps({"iterations": n, **gan.trainer.losses()}))
    else:
        gan.trainer.losses()
    gan.trainer.close()
    if distributed:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
