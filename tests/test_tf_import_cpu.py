"""Weight-file readers (SURVEY.md §8 f rank 1): TF checkpoint V2, 'checkpoint' state file, frozen GraphDef -- parsed
without TensorFlow.  Files are produced by tests/tf_format_writer.py (an independent writer of the same published
formats; no TF-written file is available offline)."""
import os

import numpy as np
import pytest

from ctpn_b200 import tf_import as T
from oracle import synth
import tf_format_writer as W


def test_crc32c_and_snappy_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                      # the standard CRC-32C check value
    assert T.crc32c(b"") == 0 and W.crc32c_bitwise(b"123456789") == 0xE3069283
    assert T.mask_crc(T.crc32c(b"123456789")) == W.masked(0xE3069283)
    # literal "abc" + copy(length 8, offset 3): overlapping copy
    assert T.snappy_decompress(bytes([11, 0x08]) + b"abc" + bytes([0x11, 0x03])) == b"abcabcabcab"
    # 2-byte-offset copy: literal "0123456789" then copy(length 5, offset 10)
    assert T.snappy_decompress(bytes([15, (10 - 1) << 2]) + b"0123456789" + bytes([((5 - 1) << 2) | 2, 10, 0])) == b"012345678901234"
    with pytest.raises(T.TFFormatError):
        T.snappy_decompress(bytes([4, 0x11, 0x09]))                  # copy before any output


def _weights_with_extras():
    w = synth.make_weights(3)
    w = {k: v for k, v in w.items()}
    w["global_step"] = np.asarray(50000, np.int64)                     # scalar int64, as a Saver stores it
    w["conv1_1/weights/Momentum"] = np.zeros_like(w["conv1_1/weights"])  # optimizer slot sharing a long key prefix
    w["half"] = np.arange(6, dtype=np.float16).reshape(2, 3)
    return w


@pytest.mark.parametrize("block_size,compress", [(4096, False), (64, False), (300, True)])
def test_checkpoint_v2_round_trip(tmp_path, block_size, compress):
    w = _weights_with_extras()
    prefix = str(tmp_path / "VGGnet_fast_rcnn_iter_50000.ckpt")
    W.write_checkpoint(prefix, w, block_size=block_size, compress=compress)
    listing = T.list_checkpoint(prefix)
    assert set(listing) == set(w)
    assert listing["conv5_3/weights"] == (np.dtype("<f4"), (3, 3, 512, 512)) and listing["global_step"][1] == ()
    got = T.read_checkpoint(prefix)
    assert set(got) == set(w)
    for k in w:
        assert got[k].dtype == w[k].dtype and got[k].shape == w[k].shape
        np.testing.assert_array_equal(got[k], w[k])
    # every accepted spelling of the location: directory (via the state file), .index, .data shard
    for loc in (str(tmp_path), prefix + ".index", prefix + ".data-00000-of-00001"):
        sub = T.read_checkpoint(loc, names=["rpn_cls_score/biases", "lstm_o/biases"])
        assert sorted(sub) == ["lstm_o/biases", "rpn_cls_score/biases"]
        np.testing.assert_array_equal(sub["lstm_o/biases"], w["lstm_o/biases"])
    with pytest.raises(KeyError):
        T.read_checkpoint(prefix, names=["no/such/variable"])


def test_checkpoint_corruption_is_detected(tmp_path):
    w = {"a/weights": np.arange(12, dtype=np.float32).reshape(3, 4), "b": np.ones(5, np.float32)}
    prefix = str(tmp_path / "m.ckpt")
    W.write_checkpoint(prefix, w)
    idx = bytearray(open(prefix + ".index", "rb").read())
    bad = bytearray(idx); bad[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(T.TFFormatError):
        T.read_checkpoint(prefix)
    bad = bytearray(idx); bad[-1] ^= 0xFF                                  # magic number
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(T.TFFormatError):
        T.read_checkpoint(prefix)
    open(prefix + ".index", "wb").write(bytes(idx))
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[3] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(T.TFFormatError):
        T.read_checkpoint(prefix)                                           # tensor crc
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data[:20]))
    with pytest.raises(T.TFFormatError):
        T.read_checkpoint(prefix, verify=False)                             # truncated shard


def test_large_tensor_corruption_is_detected(tmp_path):
    """Every tensor is checksummed (conv / LSTM weights are far above the old 64 KB limit): one flipped bit in
    the middle of a 2.4 MB conv kernel must not load."""
    w = {"conv4_2/weights": np.random.RandomState(0).standard_normal((3, 3, 256, 256)).astype(np.float32), "b": np.ones(5, np.float32)}
    prefix = str(tmp_path / "big.ckpt")
    W.write_checkpoint(prefix, w)
    got = T.read_checkpoint(prefix)
    np.testing.assert_array_equal(got["conv4_2/weights"], w["conv4_2/weights"])
    path = prefix + ".data-00000-of-00001"
    data = bytearray(open(path, "rb").read())
    data[len(data) // 2] ^= 0x10
    open(path, "wb").write(bytes(data))
    with pytest.raises(T.TFFormatError, match="checksum"):
        T.read_checkpoint(prefix)
    assert T.crc32c_fast(b"123456789") == 0xE3069283


def test_latest_checkpoint_state_file(tmp_path):
    assert T.latest_checkpoint(str(tmp_path)) is None
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "VGGnet_fast_rcnn_iter_50000.ckpt"\n'
                                         'all_model_checkpoint_paths: "VGGnet_fast_rcnn_iter_40000.ckpt"\n')
    assert T.latest_checkpoint(str(tmp_path)) == str(tmp_path / "VGGnet_fast_rcnn_iter_50000.ckpt")
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "/abs/path/model.ckpt"\n')
    assert T.latest_checkpoint(str(tmp_path)) == "/abs/path/model.ckpt"


def test_frozen_graph_constants(tmp_path):
    w = synth.make_weights(5)
    w["three"] = np.asarray([1.5, -2.0, 3.25], np.float32)                  # stored as packed float_val by the writer
    path = str(tmp_path / "ctpn.pb")
    W.write_frozen_graph(path, w, scalar_fill={"fill": ((2, 3), 0.5)})
    got = T.read_frozen_graph(path)
    assert set(got) == set(w) | {"fill"}                                    # Identity / Placeholder nodes are not constants
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
    np.testing.assert_array_equal(got["fill"], np.full((2, 3), 0.5, np.float32))
    sub = T.read_frozen_graph(path, names=["conv1_1/weights"])
    assert list(sub) == ["conv1_1/weights"]
    with pytest.raises(KeyError):
        T.read_frozen_graph(path, names=["conv9_9/weights"])


def test_load_weight_file_dispatch(tmp_path):
    """Engine.load_weight_file / Session.restore accept what the reference restores from: a checkpoint prefix or directory
    (demo.py:88-90), a frozen .pb (demo_pb.py), the VGG .npy dict (network.py:40-53) and this repo's .npz."""
    from ctpn_b200.engine import load_weight_file, REQUIRED_VARIABLES
    w = synth.make_weights(1)
    assert sorted(REQUIRED_VARIABLES) == sorted(w)
    extra = dict(w)
    extra["global_step"] = np.asarray(7, np.int64)
    prefix = str(tmp_path / "model.ckpt")
    W.write_checkpoint(prefix, extra)
    W.write_frozen_graph(str(tmp_path / "g.pb"), w)
    np.savez(str(tmp_path / "w.npz"), **w)
    for loc in (prefix, str(tmp_path), str(tmp_path / "g.pb"), str(tmp_path / "w.npz")):
        got = load_weight_file(loc)
        assert sorted(got) == sorted(w), loc                                # exactly the 38 network variables
        for k in w:
            np.testing.assert_array_equal(got[k], w[k])
    os.remove(prefix + ".index")
    with pytest.raises((FileNotFoundError, OSError)):
        load_weight_file(prefix)


def test_command_line_list_and_convert(tmp_path, capsys):
    w = {"a/weights": np.arange(6, dtype=np.float32).reshape(2, 3), "global_step": np.asarray(3, np.int64)}
    prefix = str(tmp_path / "m.ckpt")
    W.write_checkpoint(prefix, w)
    assert T._main([prefix]) == 0
    out = capsys.readouterr().out
    assert "a/weights" in out and "(2, 3)" in out and "global_step" in out
    assert T._main([str(tmp_path), str(tmp_path / "w.npz")]) == 0
    with np.load(str(tmp_path / "w.npz")) as z:
        np.testing.assert_array_equal(z["a/weights"], w["a/weights"])
