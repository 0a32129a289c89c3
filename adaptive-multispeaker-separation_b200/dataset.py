"""Data layer of the hot path's caller: the reference's TFRecord speaker files -> batches of
(mix, non_mix, ind) in the input contract of models/network.py:44-88.

Reference: data/dataset.py
  * :398-442  from_flac_to_tfrecords -- one TFRecord file per (split, sex): tf.train.Example{'audio': float32 bytes, 'key': int64}
  * :444-455  decode
  * :456-460  normalize  (tf.nn.moments over the utterance, population variance; (mean, var) kept for post-processing)
  * :462-468  mix        (mixture = sum of the stacked sources)
  * :470-499  is_long_enough / filtering (all speaker keys of a mixture distinct) / chunk (floor(L / chunk) full chunks)
  * :501-518  process    (zip the 2^N sex combinations, interleave, batch)
  * :520-645  TFDataset  (per-speaker streams: decode -> normalize -> shuffle(100) -> filter -> chunk -> unbatch -> shuffle(10);
                          zip N streams -> filtering -> mix -> batch -> prefetch(1); train / valid / test / test_other splits)

Everything here is host-side Python / numpy (the reference's is a tf.data graph on the host): the TFRecord container and the
tf.train.Example protobuf are read and written directly (no TensorFlow: record = u64 length, masked CRC-32C of the length,
payload, masked CRC-32C of the payload), so files written by the reference's from_flac_to_tfrecords load unchanged.  The
device side of the contract (mixture built from the sources, --dataset_normalize on whole chunks) stays in
Trainer.prepare / amss_prepare_inputs; a batch leaves this module as pinned-host-ready numpy arrays and `mix` is None unless
`host_mix=True` (the sources alone carry the information; the host ships a third less).

Deviation, stated: tf.data's shuffle draws from TF's own Philox stream; the buffered shuffle here has the same buffer sizes
and semantics (fill the buffer, emit a uniformly drawn element, refill) on a numpy RandomState seeded like the reference
(seed = speaker-stream index), so the ORDER of mixtures differs from a TF run while their distribution does not.
"""
import itertools
import os
import struct

import numpy as np

from .tf_bundle import crc32c, mask_crc, _get_varint, _put_varint

SPLITS = ("train", "valid", "test", "test_other")


# ------------------------------------------------------------------------------------------------------------------
# TFRecord container
# ------------------------------------------------------------------------------------------------------------------
class TFRecordWriter:
    """tf.python_io.TFRecordWriter (data/dataset.py:423, 440)."""

    def __init__(self, path):
        self.f = open(path, "wb")

    def write(self, payload):
        head = struct.pack("<Q", len(payload))
        self.f.write(head + struct.pack("<I", mask_crc(crc32c(head))) + payload + struct.pack("<I", mask_crc(crc32c(payload))))

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def tfrecord_iter(path, verify=True):
    """Payloads of a TFRecord file (tf.data.TFRecordDataset, data/dataset.py:524); CRCs are checked unless verify=False."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError(f"{path}: truncated record header")
            (n,), (hcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and mask_crc(crc32c(head[:8])) != hcrc:
                raise IOError(f"{path}: corrupted record length")
            body = f.read(n + 4)
            if len(body) < n + 4:
                raise IOError(f"{path}: truncated record")
            if verify and mask_crc(crc32c(body[:n])) != struct.unpack("<I", body[n:])[0]:
                raise IOError(f"{path}: corrupted record payload")
            yield body[:n]


# ------------------------------------------------------------------------------------------------------------------
# tf.train.Example  (Example{1: Features{1: map<string, Feature{1: BytesList{1: bytes}, 2: FloatList, 3: Int64List{1: varint}}>}})
# ------------------------------------------------------------------------------------------------------------------
def _ld(field, payload):          # length-delimited field
    return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


def encode_example(audio, key):
    """The record from_flac_to_tfrecords writes (data/dataset.py:431-437): 'audio' = float32 samples as bytes, 'key' = speaker id."""
    audio = np.ascontiguousarray(audio, dtype=np.float32).tobytes()
    f_audio = _ld(1, _ld(1, audio))                                                   # Feature.bytes_list.value
    f_key = _ld(3, _ld(1, _put_varint(int(key) & 0xFFFFFFFFFFFFFFFF)))                 # Feature.int64_list.value ([packed = true])
    entries = b"".join(_ld(1, _ld(1, name) + _ld(2, feat)) for name, feat in ((b"audio", f_audio), (b"key", f_key)))
    return _ld(1, entries)


def _fields(buf):
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _get_varint(buf, pos)
        elif wire == 2:
            n, pos = _get_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wire == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        elif wire == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wire}")
        yield field, wire, v


def decode_example(payload):
    """decode (data/dataset.py:444-455): -> (audio float32 [L], key int).  Accepts packed and unpacked int64 lists."""
    audio, key = None, None
    for f, _, features in _fields(payload):
        if f != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            name, feat = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    name = bytes(v)
                elif f3 == 2:
                    feat = v
            for f4, _, lst in _fields(feat):
                if name == b"audio" and f4 == 1:
                    for f5, _, v in _fields(lst):
                        if f5 == 1:
                            audio = np.frombuffer(bytes(v), dtype=np.float32)
                elif name == b"key" and f4 == 3:
                    for f5, wire, v in _fields(lst):
                        if f5 == 1:
                            key = v if wire == 0 else _get_varint(v, 0)[0]
    if audio is None or key is None:
        raise ValueError("record is not an {'audio', 'key'} example (data/dataset.py:447-450)")
    if key >= 1 << 63:
        key -= 1 << 64
    return audio, int(key)


def write_speaker_file(path, utterances):
    """[(audio, key), ...] -> one '<split>_<sex>.tfrecords' file in the reference's format."""
    with TFRecordWriter(path) as w:
        for audio, key in utterances:
            w.write(encode_example(audio, key))


# ------------------------------------------------------------------------------------------------------------------
# the tf.data stages
# ------------------------------------------------------------------------------------------------------------------
def buffered_shuffle(it, buffer_size, rng):
    """tf.data.Dataset.shuffle(buffer_size): fill a buffer, emit a uniformly drawn slot, refill it from the stream."""
    buf = []
    for x in it:
        if len(buf) < buffer_size:
            buf.append(x)
            continue
        i = rng.randint(buffer_size)
        out, buf[i] = buf[i], x
        yield out
    while buf:
        yield buf.pop(rng.randint(len(buf)))


def normalize(audio):
    """normalize (data/dataset.py:456-460): (x - mean) / sqrt(var), population variance; returns the (mean, var) kept for
    post-processing."""
    a = audio.astype(np.float32)
    mean = a.mean(dtype=np.float64)
    var = a.var(dtype=np.float64)
    return ((a - np.float32(mean)) / np.float32(np.sqrt(var))).astype(np.float32), (np.float32(mean), np.float32(var))


def speaker_stream(path, chunk_size, seed, dataset_normalize=False, verify=True):
    """TFDataset.get_data (data/dataset.py:523-531): chunks [chunk_size] of one (split, sex) file with their speaker key
    (and, with --dataset_normalize, the utterance's (mean, var))."""
    rng = np.random.RandomState(seed)

    def utterances():
        for payload in tfrecord_iter(path, verify):
            audio, key = decode_example(payload)
            if dataset_normalize:
                audio, st = normalize(audio)
                yield audio, key, st
            else:
                yield audio, key, None

    def chunks():
        for audio, key, st in buffered_shuffle(utterances(), 100, rng):
            if not chunk_size < audio.shape[0]:                      # is_long_enough: strictly longer than a chunk (:470-471)
                continue
            for i in range(audio.shape[0] // chunk_size):            # chunk: the tail is dropped (:483-494)
                yield audio[i * chunk_size:(i + 1) * chunk_size], key, st

    return buffered_shuffle(chunks(), 10, rng)


def mix_streams(streams):
    """zip -> filtering (every key of a mixture distinct, :473-481) -> (non_mix [S, L], ind [S], stats)."""
    for items in zip(*streams):
        keys = [k for _, k, _ in items]
        if len(set(keys)) != len(keys):
            continue
        st = None if items[0][2] is None else np.array([it[2] for it in items], dtype=np.float32)
        yield np.stack([a for a, _, _ in items]), np.array(keys, dtype=np.int64), st


def batches(mixtures, batch_size, host_mix=False, drop_remainder=False):
    """batch (:513, 596-613): (mix [B, L] or None, non_mix [B, S, L] float32, ind [B, S] int64[, meanstd [B, S, 2]])."""
    nm, ind, st = [], [], []

    def emit():
        non_mix, I = np.stack(nm), np.stack(ind)
        mix = non_mix.sum(1, dtype=np.float32) if host_mix else None     # mix (:462-468): reduce_sum over the stacked sources
        out = (mix, non_mix, I) + ((np.stack(st),) if st and st[0] is not None else ())
        nm.clear(); ind.clear(); st.clear()
        return out

    for a, k, s in mixtures:
        nm.append(a); ind.append(k); st.append(s)
        if len(nm) == batch_size:
            yield emit()
    if nm and not drop_remainder:
        yield emit()


class TFDataset:
    """TFDataset (data/dataset.py:520-645) over '<split>_<M|F>.tfrecords' files in `workdir`.

    kwargs follow the reference: batch_size, chunk_size, nb_speakers, sex (subset of ['M', 'F']), no_random_picking,
    dataset_normalize.  train() / valid() / test() / test_other() each return a FRESH iterator of host batches (the role of
    the reference's initialisable iterators), which is what Trainer.train(dataset) consumes.  rank / world shard the batch
    stream over data-parallel processes (batch i goes to rank i % world): every rank reads the files, none repeats a mixture.
    With dataset_normalize=True the utterances are normalised HERE, before chunking, exactly as the reference does -- do not
    also pass --dataset_normalize to the trainer (that flag normalises the chunks it is handed on the device).
    """

    def __init__(self, workdir, batch_size, chunk_size, nb_speakers=2, sex=("M", "F"), no_random_picking=False,
                 dataset_normalize=False, host_mix=False, verify=True, rank=0, world=1, **_ignored):
        self.workdir, self.batch_size, self.chunk_size, self.N = workdir, int(batch_size), int(chunk_size), int(nb_speakers)
        self.sex = [s for s in ("M", "F") if s in sex]
        if not self.sex:
            raise ValueError("sex must contain 'M' and / or 'F' (data/dataset.py:547-559)")
        self.no_random_picking, self.dataset_normalize = bool(no_random_picking), bool(dataset_normalize)
        self.host_mix, self.verify, self.rank, self.world = host_mix, verify, int(rank), int(world)

    def _stream(self, split, s, seed):
        path = os.path.join(self.workdir, f"{split}_{s}.tfrecords")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        return speaker_stream(path, self.chunk_size, seed, self.dataset_normalize, self.verify)

    def _mixtures(self, split):
        N = self.N
        if len(self.sex) == 2 and not self.no_random_picking:
            # one zipped stream per sex combination, seeds j + N*i (:567-575); process() zips the combinations, stacks and
            # unbatches them: one mixture of every combination in turn, until the shortest combination is exhausted (:501-518)
            combos = [mix_streams([self._stream(split, s, j + N * i) for j, s in enumerate(comb)])
                      for i, comb in enumerate(itertools.product(self.sex, repeat=N))]
            for group in zip(*combos):
                yield from group
        else:
            if len(self.sex) == 2:                                   # alternate M, F, M, ... (:561-565)
                streams = [self._stream(split, self.sex[i % 2], i) for i in range(N)]
            else:                                                    # single sex (:576-586)
                streams = [self._stream(split, self.sex[0], i) for i in range(N)]
            yield from mix_streams(streams)

    def _iter(self, split):
        for i, b in enumerate(batches(self._mixtures(split), self.batch_size, self.host_mix)):
            if i % self.world == self.rank:
                # with --dataset_normalize the batch's (mean, var) per source stay on the dataset object, as the reference's
                # `meanstd` tensor does (:632-634); the step itself takes (mix, non_mix, ind)
                self.meanstd = b[3] if len(b) > 3 else None
                yield b[:3]

    def train(self):
        return self._iter("train")

    def valid(self):
        return self._iter("valid")

    def test(self):
        return self._iter("test")

    def test_other(self):
        return self._iter("test_other")
