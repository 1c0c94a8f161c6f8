"""CPU pins of the index algebra behind csrc/conv_s2.cu and hb200_conv_s2_wgrad (no GPU, no kernels): a 3x3 stride-2 pad-1
convolution equals a 2x2 stride-1 convolution over the 2x2 space-to-depth view of its input, with filter row r mapped to
(block shift ky, sub-row dy) = (0,1), (1,0), (1,1) for r = 0, 1, 2 (same for columns); the 1x1 stride-2 downsample branch is
the centre tap of a concatenated filter; the weight-gradient accumulator of the view unpacks to the 9 real taps with the
index formula of `unpack_s2_wgrad_kernel`.  Reference layers: habitat-baselines/habitat_baselines/rl/ddppo/policy/
resnet.py:26-77 (BasicBlock with downsample), :143-160 (_make_layer)."""
import torch
import torch.nn.functional as F


def s2_k(r):   # block shift of filter row / column r (csrc/conv_s2.cu::s2_k)
    return 0 if r == 0 else 1


def s2_d(r):   # sub-row / sub-column inside the 2x2 block (csrc/conv_s2.cu::s2_d)
    return 1 if r == 0 else r - 1


def space_to_depth(x):
    """NCHW [B,C,H,W] -> [B,(dy,dx,c),H/2,W/2] with channel = (dy*2+dx)*C + c (the kernels' slab order)"""
    B, C, H, W = x.shape
    v = x.view(B, C, H // 2, 2, W // 2, 2)                 # b c by dy bx dx
    return v.permute(0, 3, 5, 1, 2, 4).reshape(B, 4 * C, H // 2, W // 2)


def view_filter(w):
    """[N,C,3,3] -> [N,4C,2,2] over the view: tap (ky,kx), channel block (dy,dx) holds w[:, :, r, s]"""
    N, C = w.shape[:2]
    w2 = torch.zeros(N, 4 * C, 2, 2, dtype=w.dtype)
    for r in range(3):
        for s in range(3):
            blk = s2_d(r) * 2 + s2_d(s)
            w2[:, blk * C:(blk + 1) * C, s2_k(r), s2_k(s)] = w[:, :, r, s]
    return w2


def test_stride2_conv_is_a_2x2_conv_over_the_space_to_depth_view():
    torch.manual_seed(0)
    x = torch.randn(3, 8, 16, 12, dtype=torch.float64)
    w = torch.randn(5, 8, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=1)
    xs = F.pad(space_to_depth(x), (1, 0, 1, 0))            # one block row / column of zeros above / left only
    got = F.conv2d(xs, view_filter(w))
    torch.testing.assert_close(got, ref, rtol=1e-12, atol=1e-12)
    # exactly 9 of the 16 (tap, channel-block) pairs are populated: the MMA loop visits only those
    w2 = view_filter(torch.ones(1, 1, 3, 3))
    assert int(w2.sum()) == 9


def test_downsample_branch_is_the_centre_tap_of_the_concatenated_filter():
    torch.manual_seed(1)
    x = torch.randn(2, 4, 8, 8, dtype=torch.float64)
    wa = torch.randn(6, 4, 3, 3, dtype=torch.float64)
    wd = torch.randn(3, 4, 1, 1, dtype=torch.float64)
    wcat = torch.zeros(9, 4, 3, 3, dtype=torch.float64)
    wcat[:6] = wa
    wcat[6:, :, 1, 1] = wd[:, :, 0, 0]
    y = F.conv2d(x, wcat, stride=2, padding=1)
    torch.testing.assert_close(y[:, :6], F.conv2d(x, wa, stride=2, padding=1), rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(y[:, 6:], F.conv2d(x, wd, stride=2), rtol=1e-12, atol=1e-12)
    # and the summed data gradient of both branches is the transposed conv of the concatenated filter
    dy = torch.randn_like(y)
    dx = torch.nn.grad.conv2d_input(x.shape, wcat, dy, stride=2, padding=1)
    ref = (torch.nn.grad.conv2d_input(x.shape, wa, dy[:, :6].contiguous(), stride=2, padding=1) +
           torch.nn.grad.conv2d_input(x.shape, wd, dy[:, 6:].contiguous(), stride=2))
    torch.testing.assert_close(dx, ref, rtol=1e-12, atol=1e-12)


def test_dgrad_sub_pixel_taps():
    """conv_s2_dgrad: sub-pixel (dy,dx) of input block (by,bx) receives filter rows r = dy+1 (mod 2) from output row
    by + (dy+1-r)/2 -- restated with explicit loops against conv2d_input."""
    torch.manual_seed(2)
    B, C, N, H, W = 2, 3, 4, 8, 8
    w = torch.randn(N, C, 3, 3, dtype=torch.float64)
    dy_ = torch.randn(B, N, H // 2, W // 2, dtype=torch.float64)
    ref = torch.nn.grad.conv2d_input((B, C, H, W), w, dy_, stride=2, padding=1)
    dyp = F.pad(dy_, (0, 1, 0, 1))                          # bottom / right out-of-range = zero (TMA fill)
    dx = torch.zeros(B, C, H, W, dtype=torch.float64)
    for sy in range(2):
        for sx in range(2):
            for r in range(3):
                if (sy + 1 - r) % 2:
                    continue
                for s in range(3):
                    if (sx + 1 - s) % 2:
                        continue
                    ky, kx = (sy + 1 - r) // 2, (sx + 1 - s) // 2
                    contrib = torch.einsum("bnyx,nc->bcyx", dyp[:, :, ky:ky + H // 2, kx:kx + W // 2], w[:, :, r, s])
                    dx[:, :, sy::2, sx::2] += contrib
    torch.testing.assert_close(dx, ref, rtol=1e-12, atol=1e-12)


def test_weight_gradient_of_the_view_unpacks_to_the_nine_taps():
    torch.manual_seed(3)
    B, C, N, H, W = 2, 4, 3, 8, 12
    x = torch.randn(B, C, H, W, dtype=torch.float64)
    dy_ = torch.randn(B, N, H // 2, W // 2, dtype=torch.float64)
    ref = torch.nn.grad.conv2d_weight(x, (N, C, 3, 3), dy_, stride=2, padding=1)
    xs = F.pad(space_to_depth(x), (1, 0, 1, 0))
    dw2 = torch.nn.grad.conv2d_weight(xs, (N, 4 * C, 2, 2), dy_)          # [N, (dy,dx,c), ky, kx]
    # accumulator layout of hb200_conv_s2_wgrad: rows ((ky*2+kx)*4 + dy*2+dx)*C + c, columns n
    acc = dw2.permute(2, 3, 1, 0).reshape(16 * C, N)
    dw = torch.empty(N, C, 3, 3, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            row0 = ((s2_k(r) * 2 + s2_k(s)) * 4 + s2_d(r) * 2 + s2_d(s)) * C    # unpack_s2_wgrad_kernel
            dw[:, :, r, s] = acc[row0:row0 + C].t()
    torch.testing.assert_close(dw, ref, rtol=1e-12, atol=1e-12)


def test_engine_marks_the_stride2_block_entry_it_can_fuse():
    """host logic only (no kernels run): at config #2's 256x256 input the resnet18 engine serves layer2.0's 3x3 stride-2
    conv + 1x1 downsample with the fused kernels; layer3.0 / layer4.0 (8x8 / 4x4 outputs) stay on the gather kernels."""
    import habitat_lab_b200 as hb
    from habitat_lab_b200 import ops
    from habitat_lab_b200.rl.resnet_policy import EncoderEngine
    from habitat_lab_b200.synthetic import pointnav_spaces

    assert ops.conv_s2_supported(32, 64, 64, 32, 32) and not ops.conv_s2_supported(64, 128, 128, 16, 16)
    obs_space, act_space = pointnav_spaces(256, 256)
    pol = hb.PointNavResNetPolicy(obs_space, act_space, hidden_size=512, num_recurrent_layers=2, rnn_type="LSTM",
                                  normalize_visual_inputs=True)
    eng = EncoderEngine(pol.net.visual_encoder)
    fused = [(c.ci, c.co, c.in_hw) for convs, cd in eng.blocks for c in convs[:1] if c.s2_pair is not None]
    assert fused == [(32, 64, (32, 32))]
    pairs = [cd for convs, cd in eng.blocks if cd is not None]
    assert [cd.s2_main is not None for cd in pairs] == [True, False, False]
    assert pol._rnn_wavefront(True, 512, 2, 128) and not pol._rnn_wavefront(True, 512, 1, 128)
    assert not pol._rnn_wavefront(True, 512, 2, 1) and not pol._rnn_wavefront(False, 512, 2, 128)
