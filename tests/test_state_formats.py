"""State wire formats: flat 219,729-element vector <-> the three-tensor Triton cache layout."""
import os
import re

import numpy as np
import pytest

import tone_oracle as orc

REF = os.environ.get("TONE_REFERENCE", "/root/reference")   # Triton configs are not part of the installed package


def _random_flat(B=2, seed=0):
    rng = np.random.default_rng(seed)
    flat = rng.standard_normal((B, orc.STATE_SIZE)).astype(np.float16)
    flat[:, 80 + 2 * 30 * 384 + 16 * 384 * 30] = [7, 30][:B]        # mhsa_len
    return flat


def test_flat_split_matches_oracle_layout(tb):
    flat = _random_flat()
    a = tb.state_formats.split_flat(flat)
    b = orc.unpack_state(flat)
    for k in orc.STATE_KEYS:
        if k == "mhsa_len":
            assert (a[k].reshape(-1).astype(np.int64) == b[k].numpy()).all()
        else:
            np.testing.assert_array_equal(a[k].astype(np.float32), b[k].numpy())
    np.testing.assert_array_equal(tb.state_formats.join_flat(a), flat)


def test_triton_layout_roundtrip_and_shapes(tb):
    flat = _random_flat()
    t, c, n = tb.state_formats.flat_to_triton(flat)
    # triton/model/config.pbtxt:44-66
    assert t.shape == (2, 18, 384, 30) and t.dtype == np.float16
    assert c.shape == (2, 32, 8, 50) and c.dtype == np.float16
    assert n.shape == (2,) and n.dtype == np.int64 and n.tolist() == [7, 30]
    np.testing.assert_array_equal(tb.state_formats.triton_to_flat(t, c, n), flat)
    p = tb.state_formats.split_flat(flat)
    # mhsa is stored transposed (H, T) in front of the 16 conv caches (export.py:346-347)
    np.testing.assert_array_equal(t[:, 0], p["mhsa"][:, 0].transpose(0, 2, 1))
    np.testing.assert_array_equal(t[:, 2:], p["conv"])
    # sub2 is the head of cache_last_channel; the 1,104-element tail holds preproc | sub1 | reduction, zero padded
    np.testing.assert_array_equal(c[:, :, :, :44], p["sub2"])
    tail = c[:, :, :, 44:].reshape(2, -1)
    np.testing.assert_array_equal(tail[:, :80], p["preproc"])
    np.testing.assert_array_equal(tail[:, 80:720], p["sub1"].reshape(2, -1))
    np.testing.assert_array_equal(tail[:, 720:1104], p["reduction"].reshape(2, -1))
    assert not tail[:, 1104:].any()


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "triton", "model", "config.pbtxt")), reason="reference absent")
def test_dims_agree_with_reference_triton_config():
    txt = open(os.path.join(REF, "triton", "model", "config.pbtxt")).read()
    dims = re.findall(r'name:\s*"(cache_last_\w+?)"[^}]*?dims:\s*\[([^\]]*)\]', txt, flags=re.S)
    found = {n: [int(x) for x in d.replace(" ", "").split(",") if x] for n, d in dims}
    assert found.get("cache_last_time") == [18, 384, 30]
    assert found.get("cache_last_channel") == [32, 8, 50]


@pytest.mark.gpu
def test_engine_state_through_triton_layout(tb, weights):
    """export -> triton layout -> back -> import continues the stream identically."""
    eng = tb.Engine(weights, max_slots=4, max_batch=2)
    pcm = tb.synth.telephony_pcm(2, 2400 * 3, seed=8)
    a, b = eng.alloc_slots(2), eng.alloc_slots(2)
    for i in range(2):
        eng.step(a, pcm[:, i * 2400:(i + 1) * 2400])
    flat = eng.export_states(a)
    t, c, n = tb.state_formats.flat_to_triton(flat)
    back = tb.state_formats.triton_to_flat(t, c, n)
    np.testing.assert_array_equal(back, flat)
    for i, s in enumerate(b):
        eng.import_state(int(s), back[i])
    la, _ = eng.step(a, pcm[:, 4800:])
    lb, _ = eng.step(b, pcm[:, 4800:])
    assert np.abs(la - lb).max() < 3e-2
    eng.close()
