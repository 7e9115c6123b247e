"""TensorFlow V2 checkpoint reader (sgg_b200/tf_checkpoint.py; reference train.py:280-292 tf.train.Saver).  No TensorFlow
exists in this image, so the format restatement is pinned by the published known answers of its building blocks
(crc32c check value, leveldb crc masking, snappy element encoding, leveldb block prefix compression) and by round trips
through the writer.  PARITY UNPINNED against a real TF-written file."""
import os
import struct

import numpy as np
import pytest

from sgg_b200 import tf_checkpoint as T


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                      # the CRC-32C check value
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA                      # rfc3720 B.4 test vectors
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283       # incremental form
    assert T.mask_crc(0) == 0xA282EAD8 and T.mask_crc(0xE3069283) == (((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_snappy_elements():
    # literal "ab", then a 1-byte-offset copy (offset 2, length 6) that overlaps its own output
    assert T.snappy_decompress(bytes([8, 0x04]) + b"ab" + bytes([0x09, 0x02])) == b"abababab"
    # 2-byte-offset copy: tag = (len - 1) << 2 | 2
    assert T.snappy_decompress(bytes([7, 0x08]) + b"xyz" + bytes([(4 - 1) << 2 | 2, 3, 0])) == b"xyzxyzx"
    # long literal (length - 1 = 69 stored in one extra byte: tag 60 << 2)
    lit = bytes(range(70))
    assert T.snappy_decompress(bytes([70, 60 << 2, 69]) + lit) == lit
    with pytest.raises(ValueError):
        T.snappy_decompress(bytes([4, 0x09, 0x02]))                  # copy before any output


def test_block_prefix_compression_and_table_round_trip(tmp_path):
    keys = [f"Generator/Generator/layer_norm_basic_lstm_cell/{n}/{p}".encode() for n in ("forget", "input", "output", "state", "transform")
            for p in ("beta", "gamma")]
    entries = [(k, bytes([i]) * (i + 1)) for i, k in enumerate(sorted(keys))]
    blk = T._make_block(entries, restart_interval=4)
    assert list(T._block_entries(blk)) == entries
    shared0, _ = T._get_varint(blk, 0)
    assert shared0 == 0 and len(blk) < sum(len(k) + len(v) + 3 for k, v in entries)     # prefixes really are shared
    path = str(tmp_path / "t.index")
    T.write_table(path, entries, block_entries=3)                                       # several data blocks
    assert T.read_table(path) == entries
    raw = bytearray(open(path, "rb").read())
    raw[5] ^= 0xFF                                                                      # corrupt a data block
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_table(path)


def test_checkpoint_round_trip_and_bucket_mapping(tmp_path):
    rng = np.random.default_rng(0)
    V = 40
    names = {
        "Generator/Generator/attention_perceptron/kernel": (24 * 512 + 512, 24), "Generator/Generator/attention_perceptron/bias": (24,),
        "Generator/Generator/layer_norm_basic_lstm_cell/kernel": (1536, 2048), "Generator/Generator/decoder/kernel": (512, V),
        "Generator/Generator/decoder/bias": (V,), "Discriminator/Discriminator/decoder/kernel": (512, 1), "Discriminator/W": (V, 300),
        "Generator/Generator/conv1_1/kernel": (3, 3, 3, 64),
    }
    tensors = {k: rng.standard_normal(s).astype(np.float32) for k, s in names.items()}
    tensors["Generator/Generator/decoder/bias/Adam"] = rng.standard_normal(V).astype(np.float32)
    tensors["Generator/Generator/decoder/bias/Adam_1"] = rng.random(V).astype(np.float32)
    tensors["beta1_power"] = np.float32(0.5 ** 7)
    tensors["beta1_power_1"] = np.float32(0.5 ** 35)
    tensors["global_step"] = np.int64(12)
    prefix = str(tmp_path / "ck" / "model.ckpt")
    T.write_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    meta = T.list_variables(prefix)
    assert meta["Discriminator/W"]["shape"] == [V, 300] and meta["Discriminator/W"]["dtype"] == T.DT_FLOAT
    back = T.read_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and np.array_equal(back[k], tensors[k]), k
    only = T.read_checkpoint(prefix, names=["Discriminator/W"])
    assert list(only) == ["Discriminator/W"]
    parts = T.split_for_buckets(back)
    assert "Generator/Generator/decoder/bias" in parts["generator"] and "Discriminator/W" in parts["discriminator"]
    assert list(parts["adam_m"]) == ["Generator/Generator/decoder/bias"] and list(parts["adam_v"]) == ["Generator/Generator/decoder/bias"]
    assert parts["step"] == {"beta1_power": 7, "beta1_power_1": 35}
    assert "Generator/Generator/conv1_1/kernel" in parts["other"] and "global_step" in parts["other"]
    # a flipped byte in the data file is caught by the per-tensor crc32c
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_checkpoint(prefix)


def test_round_trip_over_random_variable_sets(tmp_path):
    """Property test (hypothesis): any set of variables with TF-style names, the four supported dtypes, 0 to 3 dimensions and
    empty tensors survives write -> sorted string table with prefix-compressed keys across several blocks -> read."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    name = st.text(alphabet="abcdefgXYZ_/0123456789", min_size=1, max_size=40)
    shape = st.lists(st.integers(0, 5), min_size=0, max_size=3)
    dtype = st.sampled_from([np.float32, np.float64, np.int32, np.int64])
    counter = [0]

    @settings(max_examples=25, deadline=None)
    @given(st.dictionaries(name, st.tuples(shape, dtype, st.integers(0, 2 ** 31 - 1)), min_size=1, max_size=90))
    def run(spec):
        counter[0] += 1
        tensors = {}
        for k, (shp, dt, seed) in spec.items():
            a = np.random.default_rng(seed).standard_normal(shp) * 100
            tensors[k] = np.asarray(a).astype(dt)
        prefix = str(tmp_path / f"p{counter[0]}" / "model.ckpt")
        T.write_checkpoint(prefix, tensors)
        back = T.read_checkpoint(prefix)
        assert sorted(back) == sorted(tensors)
        for k, v in tensors.items():
            assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
        meta = T.list_variables(prefix)
        assert all(meta[k]["shape"] == list(v.shape) for k, v in tensors.items())

    run()
