"""Generator-only inference (BASELINE configs[4]): forward pass + train:270 argmax decoding / Gumbel-max sampling
through sgg_gen_sample, checked against the CPU oracle's logits."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _sampler(B, T, V, R, chunk, seed=0):
    from sgg_b200.params import GEN, ParamBucket, make_dims
    from sgg_b200.sampling import GeneratorSampler
    from tests.util import make_problem
    prob = make_problem(B, T, V, R=R, seed=seed, dtype=torch.float64)
    bucket = ParamBucket(GEN, make_dims(B, T, V, R))
    bucket.load_state_dict({k: v.float() for k, v in prob["gp"].items()})
    return prob, GeneratorSampler(bucket, B, T, R, chunk=chunk, seed=11)


@pytest.mark.parametrize("B,T,V,R,chunk", [(8, 3, 96, 196, 0), (70, 4, 130, 196, 32), (130, 6, 600, 196, 64), (3, 30, 5000, 60, 2)])
def test_greedy_decoding_matches_oracle_argmax(B, T, V, R, chunk):
    """tokens = argmax_v logits (train:270); the logits the decoder epilogue saw equal the oracle's within 1e-3."""
    from oracle import sgg_oracle as O
    from tests.util import rel
    prob, smp = _sampler(B, T, V, R, chunk)
    ann = prob["ann_g"].to(torch.bfloat16).cuda().contiguous()
    noise = prob["noise"].float().cuda().contiguous()
    ref = O.generator_forward(prob["gp"], prob["ann_g"], prob["noise"], T)          # [B,T,V] fp64
    tokens, logits = smp.sample(ann, "greedy", noise=noise, want_logits=True)
    torch.cuda.synchronize()
    assert tokens.shape == (B, T) and logits.shape == (B, T, V)
    assert rel(logits, ref) < TOL
    # bit-exact against the same accumulators (index work: exact, lowest index on ties)
    assert torch.equal(tokens.long().cpu(), logits.argmax(dim=2).cpu())
    # against the oracle: identical wherever the oracle's top-2 gap is resolvable at the tolerance
    top2 = ref.topk(2, dim=2).values
    clear = (top2[..., 0] - top2[..., 1]) > 2e-3 * ref.abs().amax(dim=2)
    assert clear.float().mean() > 0.9
    assert torch.equal(tokens.long().cpu()[clear], ref.argmax(dim=2)[clear])
    # without the optional logits output the tokens are the same
    tokens2 = smp.sample(ann, "greedy", noise=noise).clone()
    torch.cuda.synchronize()
    assert torch.equal(tokens2, tokens)


def test_chunking_does_not_change_the_result():
    B, T, V, R = 50, 3, 200, 196
    prob, s0 = _sampler(B, T, V, R, 0)
    _, s1 = _sampler(B, T, V, R, 16)
    ann = prob["ann_g"].to(torch.bfloat16).cuda().contiguous()
    noise = prob["noise"].float().cuda().contiguous()
    t0, l0 = s0.sample(ann, "greedy", noise=noise, want_logits=True)
    t1, l1 = s1.sample(ann, "greedy", noise=noise, want_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(t0, t1)
    assert torch.allclose(l0, l1, rtol=0, atol=1e-5 * l0.abs().max().item())


def test_gemm_argmax_epilogue_ties_and_ragged_tiles():
    """The epilogue alone: lowest column wins ties, ragged last n-tile, several n-tiles merged by atomic max."""
    from sgg_b200 import ops
    g = torch.Generator().manual_seed(5)
    M, N, K = 200, 1000, 128
    A = torch.randint(-2, 3, (M, K), generator=g).float()          # small integers: products are exact
    W = torch.randint(-2, 3, (N, K), generator=g).float()
    W[777] = W[13]                                                  # exact duplicates force ties
    W[300] = W[13]
    ref = A @ W.t()
    keys = torch.zeros(M, dtype=torch.int64, device="cuda")
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(A.bfloat16().cuda(), W.bfloat16().cuda(), M, N, segs=[(0, 0, 0, 0, K)], out=out, argmax_keys=keys)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(ops.decode_argmax_keys(keys).cpu(), ref.argmax(dim=1))   # torch.argmax returns the first maximum


def test_gumbel_sampling_is_reproducible_and_follows_softmax():
    B, T, V, R = 4096, 2, 12, 20
    prob, smp = _sampler(4, T, V, R, 0)
    from sgg_b200.sampling import GeneratorSampler
    # every image identical -> identical logits; the token histogram over images estimates softmax(logits)
    ann = prob["ann_g"][:1].to(torch.bfloat16).cuda().expand(B, R, 512).contiguous()
    noise = prob["noise"][:1].float().cuda().expand(B, 512).contiguous()
    big = GeneratorSampler(smp.bucket, B, T, R, chunk=1000, seed=3)
    tok_a, logits = big.sample(ann, "gumbel", noise=noise, want_logits=True)
    tok_a = tok_a.clone()
    big._calls = 0
    tok_b = big.sample(ann, "gumbel", noise=noise).clone()       # same Philox range -> same draws
    tok_c = big.sample(ann, "gumbel", noise=noise).clone()       # next range -> different draws
    torch.cuda.synchronize()
    assert torch.equal(tok_a, tok_b)
    assert not torch.equal(tok_a, tok_c)
    # scale the logits' spread: random-init logits are O(1), so the softmax is far from one-hot
    p = torch.softmax(logits[0].double().cpu(), dim=1)            # [T,V]
    for t in range(T):
        freq = torch.bincount(tok_a[:, t].long().cpu(), minlength=V).double() / B
        sigma = (p[t] * (1 - p[t]) / B).sqrt()
        assert ((freq - p[t]).abs() <= 5 * sigma + 1e-3).all(), (freq, p[t])


def test_generator_class_sample_method():
    from sgg_b200.architectures.generator_with_attention import Generator
    g = torch.Generator().manual_seed(2)
    ann = torch.randn(6, 14, 14, 512, generator=g).bfloat16().cuda()
    noise = torch.randn(6, 512, generator=g).cuda()
    gen = Generator(150)
    logits = gen.build_generator(ann, noise=noise)
    tokens = gen.sample(ann, noise=noise)
    torch.cuda.synchronize()
    assert torch.equal(tokens.long(), logits.argmax(dim=2))


def test_greedy_tokens_match_the_committed_golden_logits():
    """tests/golden/step_B4_T3_V64_R12.json holds the oracle's generator logits for a seeded problem: the decoded
    tokens must be their argmax wherever the top-2 gap is resolvable at the tolerance."""
    import json
    import os
    from sgg_b200.params import GEN, ParamBucket, make_dims
    from sgg_b200.sampling import GeneratorSampler
    from tests.util import make_problem
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_B4_T3_V64_R12.json")) as f:
        gold = json.load(f)
    c = gold["config"]
    B, T, V, R = c["B"], c["T"], c["V"], c["R"]
    prob = make_problem(B, T, V, R=R, seed=c["seed"], dtype=torch.float64)
    bucket = ParamBucket(GEN, make_dims(B, T, V, R))
    bucket.load_state_dict({k: v.float() for k, v in prob["gp"].items()})
    smp = GeneratorSampler(bucket, B, T, R)
    tokens = smp.sample(prob["ann_g"].to(torch.bfloat16).cuda().contiguous(), "greedy",
                        noise=prob["noise"].float().cuda().contiguous()).cpu().long()
    ref = torch.tensor(gold["logits"], dtype=torch.float64).view(B, T, V)
    top2 = ref.topk(2, dim=2).values
    clear = (top2[..., 0] - top2[..., 1]) > 2e-3 * ref.abs().amax(dim=2)
    assert clear.any()
    assert torch.equal(tokens[clear], ref.argmax(dim=2)[clear])
