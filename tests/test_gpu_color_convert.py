"""GPU parity of the RGB0 -> YUV420P / NV12 conversion (video_encoder.cc:380-398) through the C
ABI: bit-exact against the oracle (libswscale's C arithmetic) and against the fixtures generated
by the real libswscale (tests/golden/make_golden_swscale.py)."""
import json
import os

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def convert(fov, mgr, rgb0, y_ls=None, c_ls=None, src_ls=None, nv12=False, fill=0xEE):
    """One frame through fov_rgb0_to_yuv420p / fov_rgb0_to_nv12 with optional padded linesizes."""
    H, W, _ = rgb0.shape
    src_ls = src_ls or 4 * W
    y_ls = y_ls or W
    cw = W if nv12 else W // 2
    c_ls = c_ls or cw
    conv = fov.VideoFrameConverter(mgr)
    src_host = np.zeros((H, src_ls), np.uint8)
    src_host[:, :4 * W] = rgb0.reshape(H, 4 * W)
    src = mgr.upload(src_host)
    y = mgr.upload(np.full((H, y_ls), fill, np.uint8))
    if nv12:
        uv = mgr.upload(np.full((H // 2, c_ls), fill, np.uint8))
        conv.RGB0ToNV12(y, y_ls, uv, c_ls, src, src_ls, W, H)
        uvh = mgr.copy_to_host(np.empty((H // 2, c_ls), np.uint8), uv)
        yh = mgr.copy_to_host(np.empty((H, y_ls), np.uint8), y)
        assert (yh[:, W:] == fill).all() and (uvh[:, W:] == fill).all()  # padding untouched
        return yh[:, :W], uvh[:, 0:W:2], uvh[:, 1:W:2]
    u = mgr.upload(np.full((H // 2, c_ls), fill, np.uint8))
    v = mgr.upload(np.full((H // 2, c_ls), fill, np.uint8))
    conv.RGB0ToYUV420P(y, y_ls, u, c_ls, v, c_ls, src, src_ls, W, H)
    yh = mgr.copy_to_host(np.empty((H, y_ls), np.uint8), y)
    uh = mgr.copy_to_host(np.empty((H // 2, c_ls), np.uint8), u)
    vh = mgr.copy_to_host(np.empty((H // 2, c_ls), np.uint8), v)
    assert (yh[:, W:] == fill).all() and (uh[:, cw:] == fill).all() and (vh[:, cw:] == fill).all()
    return yh[:, :W], uh[:, :cw], vh[:, :cw]


def test_small_case_matches_libswscale(fov, mgr):
    g = np.load(os.path.join(GOLD, "swscale_small.npz"))
    for nv12 in (False, True):
        y, u, v = convert(fov, mgr, g["rgb0"], nv12=nv12)
        assert np.array_equal(y, g["y"]) and np.array_equal(u, g["u"]) and np.array_equal(v, g["v"])
        # the reference's plain SWS_BILINEAR call on x86 (SIMD vertical scaler): <= 1 LSB in chroma
        assert np.array_equal(y, g["y_simd"])
        assert np.abs(u.astype(int) - g["u_simd"]).max() <= 1
        assert np.abs(v.astype(int) - g["v_simd"]).max() <= 1


def test_reduced_buffer_sizes_match_libswscale_hashes(fov, mgr):
    with open(os.path.join(GOLD, "swscale.json")) as fh:
        cases = json.load(fh)["cases"]
    assert any(c["W"] == 4272 for c in cases)  # the 8K reduced buffer is covered here
    for c in cases:
        frame = O.lcg_frame(c["W"], c["H"], c["seed"])
        assert O.fnv1a64(frame) == c["frame"]
        y, u, v = convert(fov, mgr, frame)
        assert (O.fnv1a64(y), O.fnv1a64(u), O.fnv1a64(v)) == (c["y"], c["u"], c["v"]), c


@pytest.mark.parametrize("W,H", [(16, 8), (18, 10), (130, 34), (258, 66), (1072, 608), (2, 8)])
def test_ragged_sizes_and_padded_linesizes_match_oracle(fov, mgr, oracle, W, H):
    rng = np.random.default_rng(W * 1000 + H)
    rgb0 = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    rgb0[..., 3] = rng.integers(0, 256, (H, W), dtype=np.uint8)  # the padding byte must not matter
    want = oracle.rgb0_to_yuv420p(rgb0)
    for nv12 in (False, True):
        for pad in (0, 1):
            cw = W if nv12 else W // 2
            got = convert(fov, mgr, rgb0, y_ls=W + 3 * pad, c_ls=cw + 5 * pad, src_ls=4 * W + 12 * pad,
                          nv12=nv12)
            for a, b, name in zip(got, want, "yuv"):
                assert np.array_equal(a, b), (name, nv12, pad)


def test_extremes_clip_like_libswscale(fov, mgr, oracle):
    W, H = 64, 16
    for value in (0, 255):
        rgb0 = np.full((H, W, 4), value, np.uint8)
        got = convert(fov, mgr, rgb0)
        want = oracle.rgb0_to_yuv420p(rgb0)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # saturated primaries and checkerboards hit the 15-bit saturation and the byte clips
    rgb0 = np.zeros((H, W, 4), np.uint8)
    rgb0[::2, ::2, 0] = 255
    rgb0[1::2, 1::2, 2] = 255
    rgb0[:, W // 2:, 1] = 255
    got = convert(fov, mgr, rgb0)
    want = oracle.rgb0_to_yuv420p(rgb0)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_batched_streams_match_single_frames(fov, mgr, oracle):
    """The serving configuration: one reduced buffer per stream, one launch."""
    n, W, H = 5, 272, 144
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (n, H, W, 4), dtype=np.uint8)
    conv = fov.VideoFrameConverter(mgr)
    src = mgr.upload(frames)
    y = mgr.Buffer(n * H * W)
    u = mgr.Buffer(n * H * W // 4)
    v = mgr.Buffer(n * H * W // 4)
    conv.RGB0ToYUV420PFrames(n, y, H * W, W, u, v, H * W // 4, W // 2, src, H * W * 4, 4 * W, W, H)
    yh = mgr.copy_to_host(np.empty((n, H, W), np.uint8), y)
    uh = mgr.copy_to_host(np.empty((n, H // 2, W // 2), np.uint8), u)
    vh = mgr.copy_to_host(np.empty((n, H // 2, W // 2), np.uint8), v)
    yn = mgr.Buffer(n * H * W)
    uv = mgr.Buffer(n * H * W // 2)
    conv.RGB0ToNV12Frames(n, yn, H * W, W, uv, H * W // 2, W, src, H * W * 4, 4 * W, W, H)
    ynh = mgr.copy_to_host(np.empty((n, H, W), np.uint8), yn)
    uvh = mgr.copy_to_host(np.empty((n, H // 2, W), np.uint8), uv)
    for f in range(n):
        wy, wu, wv = oracle.rgb0_to_yuv420p(frames[f])
        assert np.array_equal(yh[f], wy) and np.array_equal(uh[f], wu) and np.array_equal(vh[f], wv)
        assert np.array_equal(ynh[f], wy)
        assert np.array_equal(uvh[f, :, 0::2], wu) and np.array_equal(uvh[f, :, 1::2], wv)


def test_unsupported_sizes_are_rejected(fov, mgr):
    conv = fov.VideoFrameConverter(mgr)
    buf = mgr.Buffer(1 << 16)
    for W, H in [(15, 8), (16, 9), (16, 6)]:
        with pytest.raises(fov.FovError, match="even"):
            conv.RGB0ToYUV420P(buf, 64, buf, 64, buf, 64, buf, 256, W, H)
    with pytest.raises(fov.FovError, match="invalid"):
        conv.RGB0ToYUV420P(buf, 8, buf, 64, buf, 64, buf, 256, 16, 8)  # y linesize < width
    with pytest.raises(fov.FovError, match="invalid"):
        conv.RGB0ToNV12(buf, 16, buf, 8, buf, 256, 16, 8)  # uv linesize < width


def test_foveated_stream_to_encoder_surface(fov, mgr, oracle):
    """Server loop with the conversion in place of the D2H copy + sws_scale
    (video_server.cc:296-345 -> video_encoder.cc:380-398): the planes equal the oracle's conversion
    of the oracle's reduced buffer."""
    W, H = 384, 192
    ow, oh = O.reduced_size(W), O.reduced_size(H)
    frame = O.smooth_frame(W, H, seed=3)
    enc, dec, conv = fov.SATEncoder(mgr), fov.SATDecoder(mgr), fov.VideoFrameConverter(mgr)
    src = mgr.upload(frame)
    sat = mgr.Buffer(W * H * 12)
    red = mgr.upload(np.zeros((oh, ow, 4), np.uint8))
    y, u, v = mgr.Buffer(ow * oh), mgr.Buffer(ow * oh // 4), mgr.Buffer(ow * oh // 4)
    enc.EncodeFrameGPU(sat, src, W, H, 4 * W)
    dec.SampleFrameRectGPU(red, ow, oh, 4 * ow, sat, W, H, 0.4, 0.6)
    conv.RGB0ToYUV420P(y, ow, u, ow // 2, v, ow // 2, red, 4 * ow, ow, oh)
    want_red = oracle.sat_sample_rect(oracle.sat_encode(frame), ow, oh, 0.4, 0.6)
    wy, wu, wv = oracle.rgb0_to_yuv420p(want_red)
    assert np.array_equal(mgr.copy_to_host(np.empty((oh, ow), np.uint8), y), wy)
    assert np.array_equal(mgr.copy_to_host(np.empty((oh // 2, ow // 2), np.uint8), u), wu)
    assert np.array_equal(mgr.copy_to_host(np.empty((oh // 2, ow // 2), np.uint8), v), wv)


# ---- YUV420P / NV12 -> RGB0 (video_decoder.cc:165-222) ---------------------------------------

def to_rgb0(fov, mgr, y, u, v, nv12=False, pad=0, fill=0xEE):
    H, W = y.shape
    conv = fov.VideoFrameConverter(mgr)
    y_ls, c_ls, d_ls = W + 3 * pad, (W if nv12 else W // 2) + 5 * pad, 4 * W + 8 * pad

    def padded(a, ls):
        b = np.zeros((a.shape[0], ls), np.uint8)
        b[:, :a.shape[1]] = a
        return b
    dy = mgr.upload(padded(y, y_ls))
    dst = mgr.upload(np.full((H, d_ls), fill, np.uint8))
    if nv12:
        uv = np.empty((H // 2, W), np.uint8)
        uv[:, 0::2], uv[:, 1::2] = u, v
        duv = mgr.upload(padded(uv, c_ls))
        conv.NV12ToRGB0(dst, d_ls, dy, y_ls, duv, c_ls, W, H)
    else:
        du, dv = mgr.upload(padded(u, c_ls)), mgr.upload(padded(v, c_ls))
        conv.YUV420PToRGB0(dst, d_ls, dy, y_ls, du, c_ls, dv, c_ls, W, H)
    out = mgr.copy_to_host(np.empty((H, d_ls), np.uint8), dst)
    assert (out[:, 4 * W:] == fill).all()
    return out[:, :4 * W].reshape(H, W, 4)


def test_decoder_small_case_matches_libswscale(fov, mgr):
    g = np.load(os.path.join(GOLD, "swscale_small.npz"))
    for nv12 in (False, True):
        for pad in (0, 1):
            got = to_rgb0(fov, mgr, g["dec_y"], g["dec_u"], g["dec_v"], nv12=nv12, pad=pad)
            assert np.array_equal(got, g["dec_rgb0"]), (nv12, pad)


def test_decoder_frame_sizes_match_libswscale_hashes(fov, mgr):
    with open(os.path.join(GOLD, "swscale.json")) as fh:
        cases = json.load(fh)["decode_cases"]
    assert any(c["W"] == 7680 for c in cases)
    for c in cases:
        y, u, v = O.lcg_planes(c["W"], c["H"], c["seed"])
        assert (O.fnv1a64(y), O.fnv1a64(u), O.fnv1a64(v)) == (c["y"], c["u"], c["v"])
        assert O.fnv1a64(to_rgb0(fov, mgr, y, u, v)) == c["rgb0"], c
        assert O.fnv1a64(to_rgb0(fov, mgr, y, u, v, nv12=True)) == c["rgb0"], c


@pytest.mark.parametrize("W,H", [(2, 2), (6, 4), (18, 10), (130, 34), (258, 66)])
def test_decoder_ragged_sizes_match_oracle(fov, mgr, oracle, W, H):
    rng = np.random.default_rng(W * 77 + H)
    y = rng.integers(0, 256, (H, W), dtype=np.uint8)
    u = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (H // 2, W // 2), dtype=np.uint8)
    want = oracle.yuv420p_to_rgb0(y, u, v)
    for nv12 in (False, True):
        for pad in (0, 1):
            assert np.array_equal(to_rgb0(fov, mgr, y, u, v, nv12=nv12, pad=pad), want), (nv12, pad)


def test_decoder_batched_and_full_server_chain(fov, mgr, oracle):
    """NV12 surfaces -> RGB0 -> SAT -> reduced buffer -> NV12: the server loop with both swscale
    steps on the device (video_server.cc:291-345, video_encoder.cc:380-398), for n streams."""
    n, W, H = 3, 256, 128
    ow, oh = O.reduced_size(W), O.reduced_size(H)
    rng = np.random.default_rng(5)
    ys = rng.integers(0, 256, (n, H, W), dtype=np.uint8)
    uvs = rng.integers(0, 256, (n, H // 2, W), dtype=np.uint8)
    gaze = [(0.3, 0.4), (0.5, 0.5), (0.95, 0.1)]
    conv, enc, dec = fov.VideoFrameConverter(mgr), fov.SATEncoder(mgr), fov.SATDecoder(mgr)
    dy, duv = mgr.upload(ys), mgr.upload(uvs)
    rgb = mgr.Buffer(n * H * W * 4)
    sat = mgr.Buffer(n * H * W * 12)
    red = mgr.upload(np.zeros((n, oh, ow, 4), np.uint8))
    oy, ouv = mgr.Buffer(n * oh * ow), mgr.Buffer(n * oh * ow // 2)
    conv.NV12ToRGB0Frames(n, rgb, H * W * 4, 4 * W, dy, H * W, W, duv, H * W // 2, W, W, H)
    enc.EncodeFramesGPU(n, sat, H * W * 12, rgb, H * W * 4, W, H, 4 * W)
    dec.SampleFramesRectGPU(n, red, oh * ow * 4, ow, oh, 4 * ow, sat, H * W * 12, W, H, gaze)
    conv.RGB0ToNV12Frames(n, oy, oh * ow, ow, ouv, oh * ow // 2, ow, red, oh * ow * 4, 4 * ow, ow, oh)
    got_rgb = mgr.copy_to_host(np.empty((n, H, W, 4), np.uint8), rgb)
    got_y = mgr.copy_to_host(np.empty((n, oh, ow), np.uint8), oy)
    got_uv = mgr.copy_to_host(np.empty((n, oh // 2, ow), np.uint8), ouv)
    for f in range(n):
        want_rgb = oracle.yuv420p_to_rgb0(ys[f], uvs[f][:, 0::2], uvs[f][:, 1::2])
        assert np.array_equal(got_rgb[f], want_rgb)
        want_red = oracle.sat_sample_rect(oracle.sat_encode(want_rgb), ow, oh, *gaze[f])
        wy, wu, wv = oracle.rgb0_to_yuv420p(want_red)
        assert np.array_equal(got_y[f], wy)
        assert np.array_equal(got_uv[f][:, 0::2], wu) and np.array_equal(got_uv[f][:, 1::2], wv)


def test_decoder_unsupported_sizes_are_rejected(fov, mgr):
    conv = fov.VideoFrameConverter(mgr)
    buf = mgr.Buffer(1 << 16)
    for W, H in [(15, 8), (16, 9)]:
        with pytest.raises(fov.FovError, match="even"):
            conv.YUV420PToRGB0(buf, 256, buf, 64, buf, 64, buf, 64, W, H)
    with pytest.raises(fov.FovError, match="invalid"):
        conv.NV12ToRGB0(buf, 60, buf, 16, buf, 16, 16, 8)  # target linesize < 4 * width
