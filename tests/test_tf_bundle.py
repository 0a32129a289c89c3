"""TensorFlow checkpoint bundle reader / writer: format-level checks that need no TensorFlow --
the CRC-32C known answers, the LevelDB table framing (footer magic, block trailers, prefix compression across several
blocks), the BundleEntryProto fields, a round trip under the reference's variable names, and error detection."""
import importlib.util
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("amss_tf_bundle", os.path.join(ROOT, "adaptive-multispeaker-separation_b200", "tf_bundle.py"))
tfb = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tfb)


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors and the classic check value
    assert tfb.crc32c(b"123456789") == 0xE3069283
    assert tfb.crc32c(bytes(32)) == 0x8A9136AA
    assert tfb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tfb.crc32c(bytes(range(32))) == 0x46DD794E
    # LevelDB's mask is invertible and moves the value
    c = tfb.crc32c(b"foo")
    assert tfb.mask_crc(c) != c


def _variables(seed=0):
    rng = np.random.RandomState(seed)
    v = {"front/window/w": rng.randn(64).astype(np.float32), "front/bases/bases": rng.randn(64, 16).astype(np.float32),
         "prediction/W": rng.randn(1, 40, 320).astype(np.float32), "prediction/b": rng.randn(320).astype(np.float32),
         "speaker_centroids": rng.randn(251, 8).astype(np.float32), "global_epoch": np.array(3, np.int32)}
    for i in range(3):
        for d in ("forward", "backward"):
            base = f"prediction/{d}_BLSTM_{i}/rnn/basic_lstm_cell"
            v[base + "/kernel"] = rng.randn(36, 80).astype(np.float32)
            v[base + "/bias"] = rng.randn(80).astype(np.float32)
    return v


def test_round_trip_and_framing(tmp_path):
    v = _variables()
    prefix = tfb.save_checkpoint(str(tmp_path / "model-1000"), v, block_size=256)      # small blocks: several data blocks
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == 0xDB4775248B80FB57          # LevelDB table magic
    header, entries = tfb.read_index(prefix)
    assert header["num_shards"] == 1 and header["endianness"] == 0
    assert sorted(entries) == sorted(v)                                                  # keys are stored sorted
    e = entries["prediction/W"]
    assert e["dtype"] == tfb.DT_FLOAT and e["shape"] == (1, 40, 320) and e["size"] == 40 * 320 * 4 and e["shard_id"] == 0
    assert entries["global_epoch"]["shape"] == () and entries["global_epoch"]["dtype"] == tfb.DT_INT32
    got = tfb.load_checkpoint(prefix, verify=True)
    for k in v:
        assert got[k].dtype == v[k].dtype and got[k].shape == v[k].shape and np.array_equal(got[k], v[k]), k
    sub = tfb.load_checkpoint(prefix, names={"front/window/w"})
    assert list(sub) == ["front/window/w"]


def test_corruption_is_detected(tmp_path):
    v = _variables(1)
    prefix = tfb.save_checkpoint(str(tmp_path / "model-7"), v)
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tfb.read_index(prefix)
    prefix2 = tfb.save_checkpoint(str(tmp_path / "model-8"), v)
    data = bytearray(open(prefix2 + ".data-00000-of-00001", "rb").read())
    data[100] ^= 1
    open(prefix2 + ".data-00000-of-00001", "wb").write(bytes(data))
    tfb.load_checkpoint(prefix2)                                                         # unchecked read succeeds
    with pytest.raises(ValueError):
        tfb.load_checkpoint(prefix2, verify=True)
    with pytest.raises(ValueError):
        open(str(tmp_path / "junk.index"), "wb").write(b"\x00" * 100)
        tfb.read_index(str(tmp_path / "junk"))
