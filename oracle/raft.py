"""CPU oracle: RAFT optical-flow band hot path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Functional restatement (torch CPU fp32 -- on a CPU-only host the reference's `autocast()` blocks disable
themselves, SURVEY.md top) of bands/flow_raft.py:51-66,100-106 and bands/raft/{raft,corr,extractor,update}.py.
Pinned against the imported reference RAFT by oracle/tools/make_golden.py.
"""
import cv2
import numpy as np
import torch
import torch.nn.functional as F

from .da import hue_to_rgb


# --------------------------------------------------------------------------- pre-process
def raft_preprocess(frame_u8, scale=0.75):
    """flow_raft.py:100-101 + common/flow.py:13-16: cv2 cubic resize on u8, HWC->CHW float (0..255)."""
    ds = cv2.resize(frame_u8, None, fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    return torch.from_numpy(np.array(ds)).permute(2, 0, 1).float()


def cubic_resize_f64(frame_u8, scale=0.75):
    """The real-valued result of cv2.resize(frame, None, fx=fy=scale, INTER_CUBIC) before rounding, in float64
    (cv::resize with dsize empty: sampling step 1/fx on both axes, dsize = cvRound(src*fx); cv::interpolateCubic with
    A = -0.75; replicate border).  Test infrastructure: used to show that every byte where the CUDA resize differs from
    cv2 (= IPP's float pipeline in this OpenCV build) is a .5 tie of this value to within float32 rounding (~255 * 2^-23 * a few)."""
    H, W = frame_u8.shape[:2]
    h, w = int(np.rint(H * scale)), int(np.rint(W * scale))

    def axis(n_out, n_in):
        f = (np.arange(n_out, dtype=np.float64) + 0.5) * (1.0 / scale) - 0.5
        s = np.floor(f).astype(np.int64)
        x = f - s
        A = -0.75
        c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c1 = ((A + 2) * x - (A + 3)) * x * x + 1
        c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        c = np.stack([c0, c1, c2, 1.0 - c0 - c1 - c2], -1)
        return np.clip(s[:, None] + np.arange(-1, 3)[None], 0, n_in - 1), c
    ix, cx = axis(w, W)
    iy, cy = axis(h, H)
    src = frame_u8.astype(np.float64)
    hb = sum(src[:, ix[:, k], :] * cx[None, :, k, None] for k in range(4))
    return sum(hb[iy[:, k]] * cy[:, k, None, None] for k in range(4))


def input_pad(ht, wd):
    """InputPadder(mode='sintel')._pad (common/flow.py:46-53): [left, right, top, bottom]."""
    pad_ht = (((ht // 8) + 1) * 8 - ht) % 8
    pad_wd = (((wd // 8) + 1) * 8 - wd) % 8
    return [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]


# --------------------------------------------------------------------------- encoders
def _norm(x, sd, name, kind):
    if kind == "instance":  # nn.InstanceNorm2d: no affine, per-image statistics, biased var, eps 1e-5
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=1e-5)


def _resblock(x, sd, p, kind, stride):
    """ResidualBlock.forward (raft/extractor.py:47-56)."""
    y = F.relu(_norm(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=stride, padding=1), sd, p + "norm1", kind))
    y = F.relu(_norm(F.conv2d(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1), sd, p + "norm2", kind))
    if stride != 1:
        # downsample = Sequential(conv1x1 stride, norm3); state_dict key of the norm inside Sequential is downsample.1
        x = _norm(F.conv2d(x, sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"], stride=stride), sd, p + "downsample.1", kind)
    return F.relu(x + y)


def basic_encoder(x, sd, p, kind):
    """BasicEncoder.forward (raft/extractor.py:168-192)."""
    x = F.relu(_norm(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=2, padding=3), sd, p + "norm1", kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = _resblock(x, sd, f"{p}layer{li}.0.", kind, stride)
        x = _resblock(x, sd, f"{p}layer{li}.1.", kind, 1)
    return F.conv2d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"])


# --------------------------------------------------------------------------- correlation
def corr_pyramid(fmap1, fmap2, levels=4):
    """CorrBlock.__init__/corr (raft/corr.py:13-27,53-60)."""
    b, dim, ht, wd = fmap1.shape
    corr = torch.matmul(fmap1.view(b, dim, ht * wd).transpose(1, 2), fmap2.view(b, dim, ht * wd))
    corr = corr.view(b, ht, wd, 1, ht, wd) / torch.sqrt(torch.tensor(dim).float())
    corr = corr.reshape(b * ht * wd, 1, ht, wd)
    pyr = [corr]
    for _ in range(levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def corr_lookup(pyr, coords, r=4):
    """CorrBlock.__call__ (raft/corr.py:29-50) + bilinear_sampler (raft/utils/utils.py:58-72)."""
    coords = coords.permute(0, 2, 3, 1)
    b, h1, w1, _ = coords.shape
    out = []
    for i, corr in enumerate(pyr):
        d = torch.linspace(-r, r, 2 * r + 1)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)
        cl = coords.reshape(b * h1 * w1, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        H, W = corr.shape[-2:]
        xg, yg = cl.split([1, 1], dim=-1)
        grid = torch.cat([2 * xg / (W - 1) - 1, 2 * yg / (H - 1) - 1], dim=-1)
        s = F.grid_sample(corr, grid, align_corners=True)
        out.append(s.view(b, h1, w1, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# --------------------------------------------------------------------------- update block
def motion_encoder(sd, flow, corr):
    """BasicMotionEncoder.forward (raft/update.py:89-97)."""
    p = "update_block.encoder."
    cor = F.relu(F.conv2d(corr, sd[p + "convc1.weight"], sd[p + "convc1.bias"]))
    cor = F.relu(F.conv2d(cor, sd[p + "convc2.weight"], sd[p + "convc2.bias"], padding=1))
    flo = F.relu(F.conv2d(flow, sd[p + "convf1.weight"], sd[p + "convf1.bias"], padding=3))
    flo = F.relu(F.conv2d(flo, sd[p + "convf2.weight"], sd[p + "convf2.bias"], padding=1))
    out = F.relu(F.conv2d(torch.cat([cor, flo], 1), sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1))
    return torch.cat([out, flow], dim=1)


def sep_conv_gru(sd, h, x):
    """SepConvGRU.forward (raft/update.py:45-60)."""
    p = "update_block.gru."
    for s, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(F.conv2d(hx, sd[p + "convz" + s + ".weight"], sd[p + "convz" + s + ".bias"], padding=pad))
        r = torch.sigmoid(F.conv2d(hx, sd[p + "convr" + s + ".weight"], sd[p + "convr" + s + ".bias"], padding=pad))
        q = torch.tanh(F.conv2d(torch.cat([r * h, x], 1), sd[p + "convq" + s + ".weight"], sd[p + "convq" + s + ".bias"], padding=pad))
        h = (1 - z) * h + z * q
    return h


def update_block(sd, net, inp, corr, flow):
    """BasicUpdateBlock.forward (raft/update.py:127-136)."""
    mf = motion_encoder(sd, flow, corr)
    net = sep_conv_gru(sd, net, torch.cat([inp, mf], 1))
    p = "update_block.flow_head."
    d = F.conv2d(F.relu(F.conv2d(net, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    m = "update_block.mask."
    mask = 0.25 * F.conv2d(F.relu(F.conv2d(net, sd[m + "0.weight"], sd[m + "0.bias"], padding=1)), sd[m + "2.weight"], sd[m + "2.bias"])
    return net, mask, d


def upsample_flow(flow, mask):
    """RAFT.upsample_flow (raft/raft.py:73-84): convex combination over the 3x3 coarse neighbourhood."""
    N, _, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * H, 8 * W)


def coords_grid(b, ht, wd):
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(b, 1, 1, 1)


def raft_forward(sd, image1, image2, iters=12, taps=None):
    """RAFT.forward(test_mode=True) (raft/raft.py:87-146): (flow_low, flow_up)."""
    image1 = 2 * (image1 / 255.0) - 1.0
    image2 = 2 * (image2 / 255.0) - 1.0
    b = image1.shape[0]
    fm = basic_encoder(torch.cat([image1, image2], 0), sd, "fnet.", "instance")
    fmap1, fmap2 = fm[:b].float(), fm[b:].float()
    pyr = corr_pyramid(fmap1, fmap2)
    cnet = basic_encoder(image1, sd, "cnet.", "batch")
    net, inp = torch.split(cnet, [128, 128], dim=1)
    net, inp = torch.tanh(net), torch.relu(inp)
    if taps is not None:
        taps.update(fmap1=fmap1, fmap2=fmap2, net0=net, inp=inp, corr0=pyr[0], corr3=pyr[3])
    N, _, H, W = image1.shape
    coords0 = coords_grid(N, H // 8, W // 8)
    coords1 = coords_grid(N, H // 8, W // 8)
    flow_up = None
    for it in range(iters):
        corr = corr_lookup(pyr, coords1)
        if taps is not None and it == 0:
            taps["lookup0"] = corr
        flow = coords1 - coords0
        net, up_mask, delta = update_block(sd, net, inp, corr, flow)
        coords1 = coords1 + delta
        if it == iters - 1:  # test_mode only consumes the last upsample (raft.py:143-144)
            flow_up = upsample_flow(coords1 - coords0, up_mask)
        if taps is not None and it == 0:
            taps["net1"] = net
            taps["delta0"] = delta
    return coords1 - coords0, flow_up


def raft_infer(sd, image1, image2, iters=12):
    """flow_raft.infer (flow_raft.py:51-66): image1/2 [2,3,H,W] float 0..255 -> fwd, bwd flows [H,W,2] f32."""
    ht, wd = image1.shape[-2:]
    pad = input_pad(ht, wd)
    i1 = F.pad(image1, pad, mode="replicate")
    i2 = F.pad(image2, pad, mode="replicate")
    with torch.no_grad():
        _, up = raft_forward(sd, i1, i2, iters)
    H, W = up.shape[-2:]
    up = up[..., pad[2]:H - pad[3], pad[0]:W - pad[1]]
    return up[0].permute(1, 2, 0).numpy(), up[1].permute(1, 2, 0).numpy()


# --------------------------------------------------------------------------- encode
def process_flow(flow):
    """common/encode.py:113-126 (+ encode_polar :98-102, saturation :73-78): HxWx2 f32 -> (HxWx3 u8, max_dist)."""
    distances = np.sqrt(np.square(flow[..., 0]) + np.square(flow[..., 1]))
    max_distance = distances.max()
    dX = flow[..., 0] / float(max_distance)
    dY = flow[..., 1] / float(max_distance)
    rad = np.sqrt(np.square(dX) + np.square(dY))
    a = (np.arctan2(dY, dX) / np.pi + 1.0) * 0.5
    rgb = hue_to_rgb(a)
    for c in range(3):
        rgb[..., c] = rgb[..., c] * rad + (1.0 - rad)
    return (rgb * 255).astype(np.uint8), max_distance


# --------------------------------------------------------------------------- consistency masks / 16-bit flow PNG
def warp_flow(img, flow):
    """common/flow.py:19-26: sample `img` at (x + flow_x, y + flow_y) with cv2.remap bilinear, zero outside."""
    h, w = flow.shape[:2]
    flow_new = flow.copy()
    flow_new[:, :, 0] += np.arange(w)
    flow_new[:, :, 1] += np.arange(h)[:, np.newaxis]
    return cv2.remap(img, flow_new, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)


def compute_fwdbwd_mask(fwd_flow, bwd_flow, alpha_1=0.05, alpha_2=0.5):
    """common/flow.py:28-40: forward/backward consistency masks."""
    bwd2fwd = warp_flow(bwd_flow, fwd_flow)
    fwd_err = np.linalg.norm(fwd_flow + bwd2fwd, axis=-1)
    fwd_mask = fwd_err < alpha_1 * (np.linalg.norm(fwd_flow, axis=-1) + np.linalg.norm(bwd2fwd, axis=-1)) + alpha_2
    fwd2bwd = warp_flow(fwd_flow, bwd_flow)
    bwd_err = np.linalg.norm(bwd_flow + fwd2bwd, axis=-1)
    bwd_mask = bwd_err < alpha_1 * (np.linalg.norm(bwd_flow, axis=-1) + np.linalg.norm(fwd2bwd, axis=-1)) + alpha_2
    return fwd_mask, bwd_mask


def encode_flow(flow, mask):
    """common/encode.py:105-110: flow + validity mask -> HxWx3 u16 (the 16-bit PNG of --subpath_mask)."""
    flow = 2 ** 15 + flow * (2 ** 8)
    mask = mask & (np.max(flow, axis=-1) < (2 ** 16 - 1))
    mask = mask & (0 < np.min(flow, axis=-1))
    return np.concatenate([flow.astype(np.uint16), mask[..., None].astype(np.uint16) * (2 ** 16 - 1)], axis=-1)
