"""The reference's Python surface on top of the CUDA path: Generator / Discriminator classes with
their constructors, build methods, attentionMechanism and tensor attributes; SceneGraphGAN with its constructor
signature, _Generator / _Discriminator, train() and checkpoint helpers."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(B=5, T=3, V=60, R=196):
    from tests.util import make_problem
    return make_problem(B, T, V, R=R, dtype=torch.float64)


def test_generator_class_matches_reference_semantics():
    from oracle import sgg_oracle as O
    from sgg_b200.architectures.generator_with_attention import Generator
    from tests.util import rel
    B, T, V = 5, 3, 60
    prob = _problem(B, T, V)
    g = Generator(V)                                       # gen:10-11
    assert g.vocab_size == V
    ann = prob["ann_g"].reshape(B, 14, 14, 512)
    g.build_generator(ann.bfloat16(), True)               # creates the variables (reference initialisers)
    assert set(g.variables) == set(prob["gp"])            # TF variable names
    g._engine.g.load_state_dict({k: v.float() for k, v in prob["gp"].items()})
    logits = g.build_generator(ann.bfloat16(), is_training=True, noise=prob["noise"].float())
    ref, aux = O.generator_forward(prob["gp"], ann, prob["noise"], T, return_aux=True)
    assert logits.shape == (B, 3, V) and rel(logits, ref) < 1e-3                 # gen:90-91
    assert g.downsampled.shape == (B, 14, 14, 512)                               # gen:68
    assert g.flattened_context.shape == (B, 100352)                              # gen:74
    assert g.partially_flattened_context.shape == (B, 196, 512)                  # gen:75
    assert g.alpha.shape == (B, 196) and rel(g.alpha, aux["alpha"][:, -1]) < 1e-3   # gen:16, last evaluated step
    # attentionMechanism(cell_state) uses cell_state[0] = c (gen:14)
    c = torch.randn(B, 512, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    z = g.attentionMechanism((c.float().cuda(), None))
    z_ref, al_ref = O.attention_mechanism(prob["gp"], "Generator/Generator", ann.reshape(B, -1), ann.reshape(B, 196, 512), (c, None))
    assert rel(z, z_ref) < 1e-3 and rel(g.alpha, al_ref) < 1e-3
    # fresh noise per call when none is injected (gen:81)
    a1 = g.build_generator(ann.bfloat16())
    a2 = g.build_generator(ann.bfloat16())
    assert not torch.equal(a1, a2)


def test_discriminator_class_matches_reference_semantics():
    from oracle import sgg_oracle as O
    from sgg_b200.architectures.discriminator_with_attention import Discriminator
    from tests.util import rel
    B, T, V = 5, 3, 60
    prob = _problem(B, T, V)
    W = prob["dp"]["Discriminator/W"].float()
    d = Discriminator(V, W)                                # disc:9-11
    ann = prob["ann_d"].reshape(B, 14, 14, 512)
    fake = torch.randn(B, 3, V, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    d.build_discriminator(fake.float(), ann.bfloat16())
    assert set(d.variables) == set(prob["dp"])
    assert torch.equal(d.variables["Discriminator/W"].cpu(), W)     # embedding_matrix initialises Discriminator/W (train:70-72)
    d._engine.d.load_state_dict({k: v.float() for k, v in prob["dp"].items()})
    for tri in (fake, prob["real"]):                       # soft inputs and one-hot reals (disc:86-87)
        out = d.build_discriminator(tri.float(), ann.bfloat16(), is_training=False)
        ref = O.discriminator_forward(prob["dp"], tri, ann, T)
        assert out.shape == (B, 3, 1) and rel(out, ref) < 1e-3      # disc:92-93
    assert d.alpha.shape == (B, 196)
    with pytest.raises(ValueError):
        d.build_discriminator(fake.float()[:, :2], ann.bfloat16())
    with pytest.raises(ValueError):
        d.build_discriminator(fake.float(), torch.zeros(B, 221, 221, 3))   # pixels: the conv front-end is out of scope


def test_scene_graph_gan_trainer(tmp_path):
    from sgg_b200.train import SceneGraphGAN
    V, B = 50, 4
    vocab = {f"w{i}": i for i in range(V)}
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    import numpy as np
    emb = (np.random.RandomState(0).rand(V, 300).astype("float32") - 0.5) * 0.2
    np.save(tmp_path / "emb.npy", emb)
    ck, logs = str(tmp_path / "ck"), str(tmp_path / "logs")
    gan = SceneGraphGAN(ck, logs, None, str(tmp_path / "vocab.json"), str(tmp_path / "emb.npy"), None, None,
                        critic_iters=2, batch_size=B, lambda_=10, resume=False)      # train.py:23-24 argument order
    assert gan.CRITIC_ITERS == 2 and gan.BATCH_SIZE == B and gan.LAMBDA == 10
    assert gan.g.vocab_size == V and gan.d.vocab_size == V
    assert torch.allclose(gan.d.variables["Discriminator/W"].cpu(), torch.from_numpy(emb))
    n = gan.train(max_iterations=3)
    assert n == 3 and gan.trainer.iterations == 3
    lines = open(os.path.join(logs, "train_log.jsonl")).read().strip().splitlines()
    assert json.loads(lines[0])["iteration"] == 1
    ann = torch.randn(B, 14, 14, 512).bfloat16()
    fake = gan._Generator(ann)                              # train.py:85-88
    score = gan._Discriminator(fake, ann)                   # train.py:90-93
    assert fake.shape == (B, 3, V) and score.shape == (B, 3, 1)
    gan._saveModel()
    gan2 = SceneGraphGAN(ck, logs, None, str(tmp_path / "vocab.json"), str(tmp_path / "emb.npy"), None, None,
                         critic_iters=2, batch_size=B, lambda_=10, resume=True)
    for k, v in gan.g.variables.items():
        assert torch.equal(v, gan2.g.variables[k]), k
    assert int(gan2.trainer.eng.counters.item()) == 3


def test_resume_from_a_tensorflow_checkpoint(tmp_path):
    """train.py:280,288-292,348-349: --resume restores a tf.train.Saver checkpoint.  A V2 checkpoint carrying the reference's
    variable names (plus the Adam slots, beta powers and a conv-front-end variable that must be ignored) is written with
    sgg_b200.tf_checkpoint.write_checkpoint and restored into a fresh trainer through the --resume path."""
    import numpy as np
    from sgg_b200 import tf_checkpoint as T
    from sgg_b200.train import SceneGraphGAN
    V, B = 30, 4
    src = SceneGraphGAN(str(tmp_path / "a"), None, None, None, None, None, None, critic_iters=2, batch_size=B, lambda_=10,
                        resume=False, vocab_size=V, seed=4)
    src.train(max_iterations=2)
    torch.cuda.synchronize()
    e = src.trainer.eng
    tensors = {}
    for bucket in (e.g, e.d):
        mv, vv = bucket._views(bucket.m), bucket._views(bucket.v)
        for k, v in bucket.views().items():
            tensors[k] = v.cpu().numpy()
            tensors[k + "/Adam"] = mv[k].cpu().numpy()
            tensors[k + "/Adam_1"] = vv[k].cpu().numpy()
    tensors["beta1_power"] = np.float32(0.5 ** e.g.step)          # generator's optimiser is created first (train.py:258)
    tensors["beta1_power_1"] = np.float32(0.5 ** e.d.step)
    tensors["Generator/Generator/conv1_1/kernel"] = np.zeros((3, 3, 3, 64), np.float32)
    ck = tmp_path / "b"
    T.write_checkpoint(str(ck / "model.ckpt"), tensors)
    dst = SceneGraphGAN(str(ck), None, None, None, None, None, None, critic_iters=2, batch_size=B, lambda_=10, resume=True,
                        vocab_size=V, seed=99)
    d = dst.trainer.eng
    assert (d.g.step, d.d.step) == (e.g.step, e.d.step) == (2, 4)
    for a, b in ((e.g, d.g), (e.d, d.d)):
        assert torch.equal(a.theta, b.theta) and torch.equal(a.m, b.m) and torch.equal(a.v, b.v)
        assert torch.equal(a.shadow, b.shadow)                     # the bf16 operand shadow was rebuilt from the loaded weights
    assert dst.load_tf_checkpoint(str(ck / "model.ckpt")) == ["Generator/Generator/conv1_1/kernel"]


def test_recall_at_k_evaluation(tmp_path):
    """SceneGraphGAN.test (train.py:297-335): fakes ranked by critic score, R@50 / R@100 against the real triples.
    Checked against a host restatement that uses the classes' own build_generator / build_discriminator outputs."""
    from sgg_b200.train import SceneGraphGAN
    V, B, mult = 6, 16, 8            # tiny vocabulary so that generated triples do hit real ones
    gan = SceneGraphGAN(str(tmp_path / "ck"), str(tmp_path / "logs"), None, None, None, None, None,
                        critic_iters=1, batch_size=B, lambda_=10, resume=False, vocab_size=V)
    g = torch.Generator().manual_seed(9)
    batches = [(torch.randn(B, 196, 512, generator=g).bfloat16(), torch.randn(B, 196, 512, generator=g).bfloat16(),
                torch.randint(0, V, (B, 3), generator=g)) for _ in range(2)]
    out = str(tmp_path / "recalls.txt")
    r50, r100 = gan.test(batches, multiplier=mult, out_path=out)
    assert 0.0 <= r50 <= 1.0 and 0.0 <= r100 <= 1.0
    assert open(out).read() == "{}\n{}".format(r50, r100)            # train.py:333-335
    # the recall helper is the reference's set intersection (train.py:294-295)
    assert SceneGraphGAN._recall([[1, 2, 3], [1, 2, 3], [4, 5, 0]], [[1, 2, 3], [9, 9, 9]], 50.0) == 1 / 50.0
    # deterministic given the engine's RNG position: replay the same noise through the public classes
    e = gan.trainer.eng
    e._rng_off = 0
    ag, ad, lb = batches[0]
    fakes, scores = [], []
    for _ in range(mult):
        e.sample_noise()
        e.set_batch(ag.cuda().contiguous(), ad.cuda().contiguous(), lb.cuda().contiguous())
        lg = e.gen_forward().clone()
        fakes.append(lg.argmax(-1).cpu())
        scores.append(e.disc_forward(lg).mean(1).cpu())
    fake, score = torch.cat(fakes).numpy(), torch.cat(scores).numpy()
    order = (-score).argsort(kind="stable")
    want50 = SceneGraphGAN._recall(fake[order[:50]], lb.numpy(), 50.0)
    e._rng_off = 0
    got50, _ = gan.test(batches[:1], multiplier=mult, out_path=out)
    assert got50 == want50
