"""ldpc_decode_host: the host-buffer pipeline (four slots, chunk ramp) and its int8 transport.

A quantised decoder sees a channel value only through Q(x) (Main_Functions.py:321-322, 475-494) and, with VN weights,
Q(x * w) (:168-177), so the library's host threads may pack float32 chunks to int8 before they cross PCIe.  Whatever
form a chunk takes -- int8 by choice, float32 by choice, float32 because a value has no int8 form -- the results must be
those of ldpc_decode on the same float32 words, bit for bit.
"""
import os

import numpy as np
import pytest

from conftest import load_case

pytestmark = pytest.mark.gpu


def make_decoder(name, **kw):
    import ldpc_error_floor_b200 as L
    case = load_case(name)
    g = L.BaseGraph(case["proto"], case["z"], case["punct"], case["short"])
    dec = L.NMSDecoder(g, L.WeightSet(case["sharing"], dict(case["weights"])), iters=case["T"], decoding_type=2,
                       q_bit=case["q_bit"], clip_llr=case["clip"], **kw)
    return g, dec


def same(h, r):
    assert np.array_equal(h["hard_packed"].view(np.int32), r.hard_packed.cpu().numpy())
    assert np.array_equal(h["flags"], r.flags.cpu().numpy())
    assert np.array_equal(h["iters"], r.iters.cpu().numpy())
    assert np.array_equal(h["biterr"], r.biterr.cpu().numpy())


QUIRKS = np.array([0.0, -0.0, 1e-4, -1e-4, 0.25, -0.25, 0.75, 7.5, -7.5, 7.75, 8.0, 1e30, -1e30, np.inf, -np.inf, np.nan,
                   1e5, -1e5, 2e5, 3.0000002, 0.24999999], dtype=np.float32)


@pytest.mark.parametrize("name, B", [("wimax_qms_333_t20", 70001), ("wimax_qms_303_t20", 33000), ("5g_r073_z72_qms_300_t20_sys", 9001),
                                     ("mackay_qms_300_t20", 250001), ("wimax_qms_qm5_323_t6", 20000), ("wimax_qms_q3_323_t6", 20000)])
def test_on_grid_words_cross_as_int8_with_identical_results(name, B):
    """Words on the quantiser grid (every word the reference's flows produce): chunks are packed, results unchanged;
    several chunks, a ragged tail, slots reused."""
    import torch
    g, dec = make_decoder(name)
    x = dec.generate(float(g.sigma([2.5])[0]), B, seed=5).reshape(B, -1)
    x = torch.clamp(x, -dec.q8_step * 127, dec.q8_step * 127)
    xh = x.cpu()
    pinned = xh.pin_memory()
    for et in (False, True):
        r = dec.decode(x, early_term=et)
        h = dec.decode_host(pinned, early_term=et)
        st = dec.host_stats()
        same(h, r)
        short = g.short[1] > g.short[0]   # shortened bits carry -clip_LLR: outside the int8 range when VN weights are present
        if not short:
            assert st["chunks_unencodable"] == 0, st
            assert st["chunks_q8"] >= 1 or st["threads"] < 3, st   # fewer than three host threads: pinned words stay float32
        assert st["chunks_q8"] + st["chunks_f32"] + st["chunks_unencodable"] >= 1
        h2 = dec.decode_host(xh.numpy(), early_term=et)          # pageable: every chunk packed
        st2 = dec.host_stats()
        same(h2, r)
        if not short:
            assert st2["chunks_f32"] == 0 and st2["chunks_unencodable"] == 0, st2
            assert st2["h2d_bytes"] == B * g.NZ, st2


def test_off_grid_words_without_vn_weights_are_quantised_on_the_host():
    """No VN weights: Q(x) is all the decoder uses, so raw float32 LLRs -- including the quirk values -- are packed too."""
    import torch
    g, dec = make_decoder("5g_r073_z72_qms_300_t20_sys")
    B = 6000
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((B, g.NZ)) * 3.0 + 2.0).astype(np.float32)
    x[:, :QUIRKS.size] = QUIRKS
    x[5] = np.tile(QUIRKS, g.NZ // QUIRKS.size + 1)[:g.NZ]
    xd = torch.from_numpy(x).cuda()
    for et in (False, True):
        r = dec.decode(xd, early_term=et)
        h = dec.decode_host(x, early_term=et)
        st = dec.host_stats()
        same(h, r)
        assert st["chunks_q8"] >= 1 and st["chunks_unencodable"] == 0 and st["chunks_f32"] == 0, st
    q, bad = dec.pack_q8(x)
    assert bad == 0
    rq = dec.decode_q8(torch.from_numpy(q).cuda())
    r = dec.decode(xd)
    assert torch.equal(rq.hard_packed, r.hard_packed) and torch.equal(rq.flags, r.flags)


def test_off_grid_words_with_vn_weights_stay_float32():
    """VN weights form Q(x * w) from the raw value: off-grid words have no int8 form, the chunk travels as float32."""
    import torch
    g, dec = make_decoder("wimax_qms_333_t20")
    B = 40000
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((B, g.NZ)) * 3.0 + 2.0).astype(np.float32)
    x[:, :QUIRKS.size] = QUIRKS
    xd = torch.from_numpy(x).cuda()
    r = dec.decode(xd)
    for src in (x, torch.from_numpy(x).pin_memory()):
        h = dec.decode_host(src)
        st = dec.host_stats()
        same(h, r)
        assert st["chunks_q8"] == 0 and (st["chunks_unencodable"] >= 1 or st["threads"] < 3), st   # < 3 host threads: pinned words are not even tried
    # a word set that is on the grid except for one value in the last chunk
    y = dec.generate(float(g.sigma([2.5])[0]), B, seed=9).reshape(B, -1).cpu().numpy().copy()
    y[B - 3, 17] = 0.3
    r = dec.decode(torch.from_numpy(y).cuda())
    h = dec.decode_host(y)
    st = dec.host_stats()
    same(h, r)
    assert st["chunks_unencodable"] == 1 and st["chunks_q8"] >= 1, st
    _, bad = dec.pack_q8(y)
    assert bad == 1


def test_switches_and_decoders_without_an_int8_form():
    import torch
    import ldpc_error_floor_b200 as L
    g, dec = make_decoder("wimax_qms_333_t20")
    B = 30000
    x = dec.generate(float(g.sigma([3.0])[0]), B, seed=2).reshape(B, -1)
    r = dec.decode(x)
    os.environ["LDPC_B200_NO_HOST_PACK"] = "1"
    try:
        h = dec.decode_host(x.cpu().numpy())
        st = dec.host_stats()
    finally:
        del os.environ["LDPC_B200_NO_HOST_PACK"]
    same(h, r)
    assert st["chunks_q8"] == 0 and st["h2d_bytes"] == B * g.NZ * 4, st
    # q_bit 6 saturates at 15.5, off its own grid; float decoders have no grid
    g6, dec6 = make_decoder("wimax_qms_q6_323_t6")
    x6 = dec6.generate(float(g6.sigma([3.0])[0]), 5000, seed=2).reshape(5000, -1)
    same(dec6.decode_host(x6.cpu().numpy()), dec6.decode(x6))
    assert dec6.host_stats()["chunks_q8"] == 0
    with pytest.raises(Exception):
        dec6.pack_q8(x6.cpu().numpy())
    # APP output keeps the float32 words
    h = dec.decode_host(x[:300].cpu().numpy(), app="last")
    ra = dec.decode(x[:300], app="last")
    assert np.array_equal(h["app"], ra.app.reshape(300, -1).cpu().numpy())
    # empty and tiny batches
    for n in (0, 1, 7):
        h = dec.decode_host(x[:n].cpu().numpy())
        assert h["flags"].shape == (n,)
        if n:
            same(h, dec.decode(x[:n]))
