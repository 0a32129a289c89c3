"""Data layer (data/dataset.py:398-645): TFRecord container, tf.train.Example codec and the TFDataset pipeline.
CPU only.  The record / protobuf codecs are checked against independent implementations (the `protobuf` runtime with the
tensorflow Example schema rebuilt from its published .proto, and a known-answer CRC), the pipeline against its invariants."""
import os
import struct

import numpy as np
import pytest

import amss_b200  # noqa: F401
from amss_b200 import dataset as D
from amss_b200.tf_bundle import crc32c, mask_crc


def _example_classes():
    """tensorflow/core/example/{feature,example}.proto rebuilt with the protobuf runtime (no TensorFlow needed)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="tf_example_test.proto", package="tensorflow", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname, packed in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
            if packed:
                f.options.packed = True
        return m

    msg("BytesList", ("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None, False))
    msg("FloatList", ("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None, True))
    msg("Int64List", ("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None, True))
    feat = msg("Feature", ("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".tensorflow.BytesList", False),
               ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".tensorflow.FloatList", False),
               ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".tensorflow.Int64List", False))
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", ("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.Features.FeatureEntry", False))
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorflow.Feature")
    msg("Example", ("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".tensorflow.Features", False))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    desc = pool.FindMessageTypeByName("tensorflow.Example")
    return get(desc) if get else message_factory.MessageFactory(pool).GetPrototype(desc)


def test_example_codec_against_protobuf_runtime():
    Example = _example_classes()
    rng = np.random.RandomState(0)
    audio = rng.randn(1234).astype(np.float32)
    for key in (0, 7, 250, 40000, -3):
        ex = Example()
        ex.features.feature["audio"].bytes_list.value.append(audio.tobytes())
        ex.features.feature["key"].int64_list.value.append(key)
        a, k = D.decode_example(ex.SerializeToString())            # what TensorFlow writes -> our reader
        assert k == key and np.array_equal(a, audio)
        back = Example.FromString(D.encode_example(audio, key))      # our writer -> the protobuf runtime
        assert list(back.features.feature["key"].int64_list.value) == [key]
        assert back.features.feature["audio"].bytes_list.value[0] == audio.tobytes()
        assert set(back.features.feature.keys()) == {"audio", "key"}


def test_tfrecord_container(tmp_path):
    # known answers: CRC-32C("123456789") = 0xE3069283 (RFC 3720 B.4), TFRecord mask = rotr15 + 0xa282ead8
    assert crc32c(b"123456789") == 0xE3069283
    assert mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    path = str(tmp_path / "x.tfrecords")
    payloads = [b"", b"a", os.urandom(1000), os.urandom(70000)]
    with D.TFRecordWriter(path) as w:
        for p in payloads:
            w.write(p)
    raw = open(path, "rb").read()
    assert struct.unpack("<Q", raw[:8])[0] == 0 and len(raw) == sum(16 + len(p) for p in payloads)
    assert list(D.tfrecord_iter(path)) == payloads
    bad = bytearray(raw)
    bad[16 + 12 + 0] ^= 1                                            # flip a bit of the second record's payload
    open(path, "wb").write(bytes(bad))
    with pytest.raises(IOError):
        list(D.tfrecord_iter(path))
    assert len(list(D.tfrecord_iter(path, verify=False))) == 4
    open(path, "wb").write(raw[:-3])
    with pytest.raises(IOError):
        list(D.tfrecord_iter(path))


def _make_corpus(root, n_spk=6, chunk=400, seed=0):
    """'<split>_<sex>.tfrecords' files as from_flac_to_tfrecords writes them: every sample encodes (speaker, utterance, position)
    so a chunk can be traced back to its utterance."""
    rng = np.random.RandomState(seed)
    truth = {}
    for split in ("train", "valid", "test"):
        for si, sex in enumerate(("M", "F")):
            utts = []
            for u in range(14):
                key = si * n_spk + rng.randint(n_spk)
                L = int(rng.choice([chunk // 2, chunk, chunk + 1, 2 * chunk + 17, 3 * chunk + 5]))
                audio = (key * 1000 + u + np.arange(L) * 1e-4).astype(np.float32)
                utts.append((audio, key))
            truth[(split, sex)] = utts
            D.write_speaker_file(os.path.join(root, f"{split}_{sex}.tfrecords"), utts)
    return truth


def test_speaker_stream_filters_chunks_and_keeps_keys(tmp_path):
    chunk = 400
    truth = _make_corpus(str(tmp_path), chunk=chunk)
    got = list(D.speaker_stream(str(tmp_path / "train_M.tfrecords"), chunk, seed=3))
    want = []
    for audio, key in truth[("train", "M")]:
        if chunk < len(audio):                                       # strictly longer (is_long_enough); tail dropped (chunk)
            want += [(audio[i * chunk:(i + 1) * chunk].tobytes(), key) for i in range(len(audio) // chunk)]
    assert sorted((a.tobytes(), k) for a, k, _ in got) == sorted(want)
    assert all(a.shape == (chunk,) and st is None for a, _, st in got)
    again = list(D.speaker_stream(str(tmp_path / "train_M.tfrecords"), chunk, seed=3))
    assert [a.tobytes() for a, _, _ in again] == [a.tobytes() for a, _, _ in got]            # seeded: reproducible
    other = list(D.speaker_stream(str(tmp_path / "train_M.tfrecords"), chunk, seed=4))
    assert [a.tobytes() for a, _, _ in other] != [a.tobytes() for a, _, _ in got]            # shuffled by the seed


@pytest.mark.parametrize("sex,no_random_picking,S", [(("M", "F"), False, 2), (("M", "F"), True, 2), (("M",), False, 2),
                                                     (("M", "F"), False, 3)])
def test_tfdataset_batches_follow_the_input_contract(tmp_path, sex, no_random_picking, S):
    chunk, B = 400, 4
    _make_corpus(str(tmp_path), chunk=chunk)
    ds = D.TFDataset(str(tmp_path), batch_size=B, chunk_size=chunk, nb_speakers=S, sex=sex,
                     no_random_picking=no_random_picking, host_mix=True)
    n = 0
    for split in ("train", "valid", "test"):
        for mix, non_mix, ind in getattr(ds, split)():
            n += 1
            assert non_mix.dtype == np.float32 and ind.dtype == np.int64
            assert non_mix.shape[1:] == (S, chunk) and ind.shape == non_mix.shape[:2] and non_mix.shape[0] <= B
            assert np.array_equal(mix, non_mix.sum(1, dtype=np.float32))                       # mix = sum of the sources
            for b in range(ind.shape[0]):
                assert len(set(ind[b].tolist())) == S                                         # filtering: distinct speakers
                assert np.array_equal(np.floor(non_mix[b, :, 0] / 1000 + 1e-3).astype(np.int64), ind[b])   # key travels with its audio
            if sex == ("M",):
                assert (ind < 6).all()
            elif no_random_picking:
                assert (ind[:, 0::2] < 6).all() and (ind[:, 1::2] >= 6).all()                  # M, F, M, ... (:561-565)
    assert n > 0
    if sex == ("M", "F") and not no_random_picking and S == 2:
        # the 2^N sex combinations take turns: MM, MF, FM, FF, MM, ... (process(), :501-518)
        ind = np.concatenate([i for _, _, i in ds.train()])
        pattern = [(a >= 6, b >= 6) for a, b in ind[:8].tolist()]
        assert pattern[:4] == [(False, False), (False, True), (True, False), (True, True)] and pattern[4:8] == pattern[:4]
    ds2 = D.TFDataset(str(tmp_path), batch_size=B, chunk_size=chunk, nb_speakers=S, sex=sex, no_random_picking=no_random_picking)
    first = next(iter(ds2.train()))
    assert first[0] is None                                                                  # the device builds the mixture


def test_tfdataset_normalize_and_rank_sharding(tmp_path):
    chunk, B = 400, 2
    truth = _make_corpus(str(tmp_path), chunk=chunk)
    ds = D.TFDataset(str(tmp_path), batch_size=B, chunk_size=chunk, nb_speakers=2, sex=("M",), dataset_normalize=True)
    stats = {}
    for audio, key in truth[("train", "M")]:
        a64 = audio.astype(np.float64)
        stats[(key, int(round((a64[0] - key * 1000))))] = (a64.mean(), a64.var())
    seen = 0
    for _, non_mix, ind in ds.train():
        ms = ds.meanstd
        assert ms.shape == (non_mix.shape[0], 2, 2)
        for b in range(non_mix.shape[0]):
            for s in range(2):
                mean, var = ms[b, s]
                # undo the normalisation with the statistics the dataset kept (data/dataset.py:459-460): the utterance comes back
                x = non_mix[b, s].astype(np.float64) * np.sqrt(var) + mean
                u = int(round(x[0] - ind[b, s] * 1000))
                m_ref, v_ref = stats[(int(ind[b, s]), u)]
                assert abs(mean - m_ref) < 1e-2 * abs(m_ref) + 1e-3 and abs(var - v_ref) < 1e-3 * v_ref
                seen += 1
    assert seen > 0
    # rank sharding: the two ranks' batch streams interleave into the single-process stream
    kw = dict(batch_size=B, chunk_size=chunk, nb_speakers=2, sex=("M", "F"))
    whole = [i.tobytes() for _, _, i in D.TFDataset(str(tmp_path), **kw).train()]
    r0 = [i.tobytes() for _, _, i in D.TFDataset(str(tmp_path), rank=0, world=2, **kw).train()]
    r1 = [i.tobytes() for _, _, i in D.TFDataset(str(tmp_path), rank=1, world=2, **kw).train()]
    assert r0 == whole[0::2] and r1 == whole[1::2] and len(whole) > 2
