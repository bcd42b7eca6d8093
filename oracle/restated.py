"""TEST INFRASTRUCTURE — plain-torch CPU restatement of the reference's SR hot path.

Functional (no nn.Module): every function takes the reference's own state_dict (same keys,
shapes, layouts) and follows the cited reference lines.  All arithmetic is torch CPU ops in the
dtype of the inputs (fp32, or fp64 for the high-precision oracle).  The algorithmic content of
`torch.nn.Conv2d / ConvTranspose2d / PReLU / PixelShuffle` lives in PyTorch (pinned by the
reference to pytorch=1.1.0, env.yml:150; torch 2.11 here — semantics unchanged for these ops).

Pinned against golden vectors produced by the real reference (oracle/make_golden.py ->
tests/golden/, checked in tests/test_oracle.py).
"""
import math

import torch
import torch.nn.functional as F

# (kernel, stride, padding) of the projection pair per upscale factor — drf_net.py:70-77
PROJ = {2: (6, 2, 2), 3: (7, 3, 2), 4: (8, 4, 2), 8: (12, 8, 2)}


def _prelu(x, sd, key):
    # nn.PReLU(num_parameters=1) — drf_net.py:56,58,66,...
    return F.prelu(x, sd[key + ".weight"])


def _conv(x, sd, key, **kw):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], **kw)


def _deconv(x, sd, key, **kw):
    return F.conv_transpose2d(x, sd[key + ".weight"], sd[key + ".bias"], **kw)


def num_groups_of(sd):
    g = 0
    while f"f_block.down_blocks.{g}.prelu.weight" in sd or f"f_block.down_blocks.{g}.prelu2.weight" in sd:
        g += 1
    return g


def in_block(x, sd, p="in_block"):
    # _InBlock — drf_net.py:52-58: conv3x3(p=1) + PReLU, conv1x1 + PReLU
    x = _prelu(_conv(x, sd, f"{p}.conv1", padding=1), sd, f"{p}.prelu1")
    return _prelu(_conv(x, sd, f"{p}.conv2"), sd, f"{p}.prelu2")


def f_block(x, hidden, sd, upscale, p="f_block"):
    # _FBlock.forward — drf_net.py:118-133
    k, s, pad = PROJ[upscale]
    G = num_groups_of(sd) if p == "f_block" else None
    feats = torch.cat([x, hidden], dim=1)                                            # :119
    lr = _prelu(_conv(feats, sd, f"{p}.in_block.conv"), sd, f"{p}.in_block.prelu")   # :120
    lr_list, hr_list = [lr], []
    g = 0
    while True:
        if f"{p}.up_blocks.{g}.deconv.weight" in sd:                                 # i == 0, :79-88
            hr = _prelu(_deconv(torch.cat(lr_list, 1), sd, f"{p}.up_blocks.{g}.deconv", stride=s, padding=pad),
                        sd, f"{p}.up_blocks.{g}.prelu")
            hr_list.append(hr)
            lr = _prelu(_conv(torch.cat(hr_list, 1), sd, f"{p}.down_blocks.{g}.conv", stride=s, padding=pad),
                        sd, f"{p}.down_blocks.{g}.prelu")
        elif f"{p}.up_blocks.{g}.conv1.weight" in sd:                                # i > 0, :89-102
            t = _prelu(_conv(torch.cat(lr_list, 1), sd, f"{p}.up_blocks.{g}.conv1"), sd, f"{p}.up_blocks.{g}.prelu1")
            hr = _prelu(_deconv(t, sd, f"{p}.up_blocks.{g}.deconv2", stride=s, padding=pad),
                        sd, f"{p}.up_blocks.{g}.prelu2")
            hr_list.append(hr)
            t = _prelu(_conv(torch.cat(hr_list, 1), sd, f"{p}.down_blocks.{g}.conv1"), sd, f"{p}.down_blocks.{g}.prelu1")
            lr = _prelu(_conv(t, sd, f"{p}.down_blocks.{g}.conv2", stride=s, padding=pad),
                        sd, f"{p}.down_blocks.{g}.prelu2")
        else:
            break
        lr_list.append(lr)
        g += 1
    del G
    feats = torch.cat(lr_list[1:], dim=1)                                            # :131
    return _prelu(_conv(feats, sd, f"{p}.out_block.conv"), sd, f"{p}.out_block.prelu")  # :132


def out_block(x, sd, upscale, p="out_block"):
    # _OutBlock — drf_net.py:136-147
    if math.log(upscale, 2) % 1 == 0:
        n = int(math.log(upscale, 2))
        for i in range(n):
            x = F.pixel_shuffle(_conv(x, sd, f"{p}.conv{i + 1}", padding=1), 2)      # :141-142
        return _conv(x, sd, f"{p}.conv{n + 1}", padding=1)                           # :144
    x = F.pixel_shuffle(_conv(x, sd, f"{p}.conv1", padding=1), 3)                    # :145-146
    return _conv(x, sd, f"{p}.conv2", padding=1)                                     # :147


def drfnet_forward(inputs, sd, upscale):
    """DRFNet.forward — drf_net.py:38-49. inputs: list of T tensors [N,C,h,w]."""
    outputs, hidden = [], None
    for i, x in enumerate(inputs):
        feats = in_block(x, sd)                     # :41
        if i == 0:
            hidden = feats                          # :42-43
        f = f_block(feats, hidden, sd, upscale)     # :44
        hidden = f                                  # :45
        outputs.append(out_block(feats + f, sd, upscale))  # :46-48
    return outputs


def srfbnet_forward(x, sd, upscale, num_steps):
    """SRFBNet.forward — srfb_net.py:38-50 (lrf_block = _InBlock layout, _RBlock :137-151)."""
    k, s, pad = PROJ[upscale]
    outputs, hidden = [], None
    for i in range(num_steps):
        feats = in_block(x, sd, p="lrf_block")                                       # :41
        if i == 0:
            hidden = feats                                                           # :42-43
        feats = f_block(feats, hidden, sd, upscale)                                  # :44
        hidden = feats                                                               # :45
        r = _prelu(_deconv(feats, sd, "r_block.deconv1", stride=s, padding=pad), sd, "r_block.prelu1")
        r = _conv(r, sd, "r_block.conv2", padding=1)                                 # :46
        up = F.interpolate(x, scale_factor=upscale, mode="bilinear", align_corners=False)   # :47
        outputs.append(up + r)                                                       # :48
    return outputs


# ---- losses -----------------------------------------------------------------------------
def l1_loss(o, t):          # torch.nn.L1Loss resolved by name, main.py:60-63
    return (o - t).abs().mean()


def mse_loss(o, t):         # torch.nn.MSELoss
    return ((o - t) ** 2).mean()


def charbonnier_loss(o, t, epsilon):   # losses.py:23-34
    return torch.sqrt((o - t) ** 2 + epsilon).mean()


def huber_loss(o, t, delta):           # losses.py:5-20
    abs_error = (o - t).abs()
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return (0.5 * quadratic ** 2 + delta * linear).mean()


LOSSES = {"L1Loss": (0, l1_loss), "MSELoss": (1, mse_loss),
          "CharbonnierLoss": (2, charbonnier_loss), "HuberLoss": (3, huber_loss)}


def vsr_loss(outputs, targets, fn):
    # AcdcVSRTrainer._compute_losses — acdc_vsr_trainer.py:74-88: mean over frames of per-frame loss
    return torch.stack([fn(o, t) for o, t in zip(outputs, targets)]).mean()


# ---- metrics ----------------------------------------------------------------------------
DATASET_STATS = {"acdc": (54.089, 48.084), "dsb15": (51.193, 52.671)}   # utils.py:13-16


def denormalize(x, dataset):   # utils.py:1-20
    mean, std = DATASET_STATS[dataset]
    return (x.clone() * std + mean).round().clamp(0, 255)


def psnr(o, t, max_value=255, size_average=True):   # metrics.py:20-36
    dims = list(range(1, o.dim()))
    mse = ((o - t) ** 2).mean(dims)
    val = 10 * torch.log10(max_value ** 2 / (mse + 1e-10))
    return val.mean() if size_average else val


def ssim_window_1d(dtype=torch.float32):
    """Per-axis factor of the reference window (metrics.py:68-77): exp(-((i-5)/(2*1.5))^2); the
    2-D kernel is the outer product, normalised to sum 1 — the 1-D factor normalised to sum 1
    gives the same outer product."""
    i = torch.arange(11, dtype=torch.float32)
    g = 1 / (1.5 * math.sqrt(2 * math.pi)) * torch.exp(-((i - 5) / (2 * 1.5)) ** 2)
    return (g / g.sum()).to(dtype)


def ssim(o, t, value_range=255, size_average=True, dim=2):   # metrics.py:51-113, channels=1
    c1, c2 = (0.01 * value_range) ** 2, (0.03 * value_range) ** 2
    grids = torch.meshgrid([torch.arange(11, dtype=torch.float32)] * dim, indexing="ij")
    k = 1
    for g in grids:
        k = k * (1 / (1.5 * math.sqrt(2 * math.pi)) * torch.exp(-((g - 5) / (2 * 1.5)) ** 2))
    k = (k / k.sum()).to(o.dtype).view(1, 1, *[11] * dim)
    conv = F.conv2d if dim == 2 else F.conv3d            # metrics.py:60-63
    mu1, mu2 = conv(o, k), conv(t, k)
    s11 = conv(o * o, k) - mu1.pow(2)
    s22 = conv(t * t, k) - mu2.pow(2)
    s12 = conv(o * t, k) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2.0 * s12 + c2)) / ((mu1.pow(2) + mu2.pow(2) + c1) * (s11 + s22 + c2))
    return m.mean() if size_average else m.mean(dim=list(range(1, o.dim())))


def cardiac(metric, o, t, box):   # CardiacPSNR / CardiacSSIM.forward — metrics.py:131-139,157-165
    h0, hn, w0, wn = box
    return metric(o[..., h0:hn, w0:wn], t[..., h0:hn, w0:wn])


def vsr_metrics(outputs, targets, dataset="acdc"):
    # AcdcVSRTrainer._compute_metrics — acdc_vsr_trainer.py:90-107
    o = [denormalize(x, dataset) for x in outputs]
    t = [denormalize(x, dataset) for x in targets]
    return (torch.stack([psnr(a, b) for a, b in zip(o, t)]).mean(),
            torch.stack([ssim(a, b) for a, b in zip(o, t)]).mean())


def edsrnet_forward(x, sd, upscale, res_scale=0.1):
    """EDSRNet.forward — edsr_net.py:34-38 with _ResBlock :41-53 and _UpBlock :56-67."""
    head = _conv(x, sd, "head.0", padding=1)                                        # :35
    y, b = head, 0
    while f"body.{b}.body.conv1.weight" in sd:
        t = _conv(torch.relu(_conv(y, sd, f"body.{b}.body.conv1", padding=1)), sd, f"body.{b}.body.conv2", padding=1)
        y = t * res_scale + y                                                        # :50-52
        b += 1
    y = _conv(y, sd, "body.conv", padding=1) + head                                 # :36
    i = 1
    while f"tail.0.conv{i}.weight" in sd:
        y = F.pixel_shuffle(_conv(y, sd, f"tail.0.conv{i}", padding=1), 3 if upscale == 3 else 2)   # :60-65
        i += 1
    return _conv(y, sd, "tail.conv", padding=1)                                     # :32,37


# ---- DUFNet (duf_net.py:9-214) ------------------------------------------------------------------------------
def _bn3d(x, sd, key, training, eps=1e-5):
    # nn.BatchNorm3d — duf_net.py:114,198,201,207,210: batch statistics (biased variance) when training
    if training:
        return F.batch_norm(x, None, None, sd[key + ".weight"], sd[key + ".bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                        False, 0.0, eps)


def dufnet_forward(inputs, sd, size_filter, upscale, training=True):
    """DUFNet.forward — duf_net.py:51-99 (any backbone: the layer count is read from the state_dict)."""
    T = len(inputs)
    t = T // 2 if T % 2 == 1 else T // 2 - 1                                                   # :53
    target = inputs[t].unsqueeze(2)
    feats = torch.stack([F.conv2d(f, sd["head.weight"], sd["head.bias"], padding=1) for f in inputs], dim=2)   # :57-61
    concat, i = feats, 0
    while f"denseLayer.conv{i}.conv2.weight" in sd:                                            # _DenseLayer*.forward :119-130
        p = f"denseLayer.conv{i}"
        x = F.relu(_bn3d(concat, sd, p + ".bn1", training))
        x = F.conv3d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
        x = F.relu(_bn3d(x, sd, p + ".bn2", training))
        valid_t = f"denseLayer.conv{i + 3}.conv2.weight" not in sd                             # the last three: _denseBlock2
        x = F.conv3d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=(0, 1, 1) if valid_t else 1)
        concat = torch.cat((concat[:, :, 1:-1] if valid_t else concat, x), dim=1)              # :125-128
        i += 1
    x = F.relu(_bn3d(concat, sd, "denseLayer.tail.bn", training))
    feats = F.conv3d(x, sd["denseLayer.tail.conv.weight"], sd["denseLayer.tail.conv.bias"], padding=(0, 1, 1))
    head2 = lambda p: F.conv3d(F.relu(F.conv3d(F.relu(feats), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])),
                               sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])                 # :37-48
    k2, rr = size_filter ** 2, upscale ** 2
    filters = head2("filterNet")
    filters = torch.softmax(filters.reshape(filters.shape[0], k2, rr, *filters.shape[2:]), dim=1)[:, :, :, 0]   # :66-72
    n, c, _, h, w = target.shape
    nb = F.unfold(target[:, :, 0].reshape(n * c, 1, h, w), size_filter, padding=size_filter // 2)   # :79-82 (identity conv)
    nb = nb.reshape(n, c, k2, h, w)
    out = torch.einsum("nckhw,nkphw->ncphw", nb, filters).reshape(n, c * rr, h, w)             # :84-88
    out = F.pixel_shuffle(out, upscale)                                                        # :89
    residual = F.pixel_shuffle(head2("residualNet").squeeze(2), upscale)                       # :93-96
    return out + residual


def drfnet_init(in_channels, out_channels, num_features, num_groups, upscale_factor):
    """state_dict of a freshly constructed DRFNet (drf_net.py:23-36,52-58,61-106,136-147): the reference's own
    torch.nn modules built in the reference's construction order, hence the reference's default initialisation
    under the current torch seed.  Lets bench.py's CPU arms run without importing the product package."""
    import collections
    import torch.nn as nn
    F_, G, r = num_features, num_groups, upscale_factor
    k, s, p = PROJ[r]
    sd = collections.OrderedDict()

    def put(prefix, mod):
        for name, v in mod.state_dict().items():
            sd[f"{prefix}.{name}"] = v.detach().clone()

    put("in_block.conv1", nn.Conv2d(in_channels, 4 * F_, 3, padding=1)); put("in_block.prelu1", nn.PReLU(1, 0.2))
    put("in_block.conv2", nn.Conv2d(4 * F_, F_, 1)); put("in_block.prelu2", nn.PReLU(1, 0.2))
    put("f_block.in_block.conv", nn.Conv2d(2 * F_, F_, 1)); put("f_block.in_block.prelu", nn.PReLU(1, 0.2))
    ups, downs = collections.OrderedDict(), collections.OrderedDict()
    for g in range(G):                                 # drf_net.py:78-102: up block then down block of every group
        u, d = f"f_block.up_blocks.{g}", f"f_block.down_blocks.{g}"
        if g == 0:
            ups[f"{u}.deconv"] = nn.ConvTranspose2d(F_, F_, k, s, p); ups[f"{u}.prelu"] = nn.PReLU(1, 0.2)
            downs[f"{d}.conv"] = nn.Conv2d(F_, F_, k, s, p); downs[f"{d}.prelu"] = nn.PReLU(1, 0.2)
        else:
            ups[f"{u}.conv1"] = nn.Conv2d(F_ * (g + 1), F_, 1); ups[f"{u}.prelu1"] = nn.PReLU(1, 0.2)
            ups[f"{u}.deconv2"] = nn.ConvTranspose2d(F_, F_, k, s, p); ups[f"{u}.prelu2"] = nn.PReLU(1, 0.2)
            downs[f"{d}.conv1"] = nn.Conv2d(F_ * (g + 1), F_, 1); downs[f"{d}.prelu1"] = nn.PReLU(1, 0.2)
            downs[f"{d}.conv2"] = nn.Conv2d(F_, F_, k, s, p); downs[f"{d}.prelu2"] = nn.PReLU(1, 0.2)
    for name, m in list(ups.items()) + list(downs.items()):      # state_dict order: all up_blocks, then all down_blocks
        put(name, m)
    put("f_block.out_block.conv", nn.Conv2d(F_ * G, F_, 1)); put("f_block.out_block.prelu", nn.PReLU(1, 0.2))
    if math.log(r, 2) % 1 == 0:
        n = int(math.log(r, 2))
        for i in range(n):
            put(f"out_block.conv{i + 1}", nn.Conv2d(F_, 4 * F_, 3, padding=1))
        put(f"out_block.conv{n + 1}", nn.Conv2d(F_, out_channels, 3, padding=1))
    else:
        put("out_block.conv1", nn.Conv2d(F_, 9 * F_, 3, padding=1))
        put("out_block.conv2", nn.Conv2d(F_, out_channels, 3, padding=1))
    return sd


# ---- RBPNet (rbp_net.py:8-285) ----------------------------------------------------------------------------------------
def rbpnet_forward(inputs, sd, upscale, num_frames):
    """RBPNet.forward - rbp_net.py:65-91 with DBPNet.forward :129-139, UpBlock :268-276, DownBlock :278-285,
    ResnetBlock.forward :227-246 (norm=None, the block's single PReLU applied twice), ConvBlock / DeconvBlock :142-219."""
    k, s, pad = PROJ[upscale]
    inputs = list(inputs)
    t = num_frames // 2 if num_frames % 2 == 1 else num_frames // 2 - 1            # :14
    x = inputs.pop(t)                                                               # :66
    act = lambda v, p: F.prelu(v, sd[p + ".act.weight"])
    cb = lambda v, p, **kw: act(_conv(v, sd, p + ".conv", **kw), p)                 # ConvBlock with PReLU
    db = lambda v, p: act(_deconv(v, sd, p + ".deconv", stride=s, padding=pad), p)  # DeconvBlock with PReLU
    proj = dict(stride=s, padding=pad)

    def resblock(v, p):
        out = act(_conv(v, sd, p + ".conv1", padding=1), p)
        return act(_conv(out, sd, p + ".conv2", padding=1) + v, p)

    def seq(v, p, last):
        i = 0
        while f"{p}.{i}.conv1.weight" in sd:
            v = resblock(v, f"{p}.{i}")
            i += 1
        return last(v, f"{p}.{i}")

    def up(v, p):
        h0 = db(v, p + ".up_conv1")
        l0 = cb(h0, p + ".up_conv2", **proj)
        return db(l0 - v, p + ".up_conv3") + h0

    def down(v, p):
        l0 = cb(v, p + ".down_conv1", **proj)
        h0 = db(l0, p + ".down_conv2")
        return cb(h0 - v, p + ".down_conv3", **proj) + l0

    def dbp(v):
        v = cb(v, "dbp_net.feat1")
        h1 = up(v, "dbp_net.up1")
        h2 = up(down(h1, "dbp_net.down1"), "dbp_net.up2")
        h3 = up(down(h2, "dbp_net.down2"), "dbp_net.up3")
        return _conv(torch.cat((h3, h2, h1), 1), sd, "dbp_net.output.conv")         # :137 (no activation)

    feat_input = cb(x, "feat0", padding=1)                                          # :70
    feat_frame = [cb(torch.cat([x, nb], dim=1), "feat1", padding=1) for nb in inputs]   # :71-73
    Ht = []
    for j in range(len(inputs)):                                                    # :77-86
        h0 = dbp(feat_input)
        h1 = seq(feat_frame[j], "res_feat1", db)
        e = seq(h0 - h1, "res_feat2", lambda v, p: cb(v, p, padding=1))
        h = h0 + e
        Ht.append(h)
        feat_input = seq(h, "res_feat3", lambda v, p: cb(v, p, **proj))
    return _conv(torch.cat(Ht, dim=1), sd, "output.conv", padding=1)                # :89-90


# ---- FRVSRNet (frvsr_net.py:11-239) -----------------------------------------------------------------------------------
def stn_mesh(h, w, dtype=torch.float32):
    """STN.nd_meshgrid(h, w, permute=[1, 0]) - frvsr_net.py:228-239: [h, w, 2] with (x, y) in linspace(-1, 1)"""
    ys = torch.linspace(-1, 1, h, dtype=torch.float64)
    xs = torch.linspace(-1, 1, w, dtype=torch.float64)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).to(dtype)


def stn_warp(img, u, v):
    """STN(mode='bilinear', padding_mode='border', normalize=False).forward - frvsr_net.py:205-226 (grid_sample with
    the framework's default align_corners, i.e. False on the torch of this image)"""
    mesh = stn_mesh(*img.shape[-2:], dtype=img.dtype).unsqueeze(0) + torch.stack([u, v], dim=-1)
    return F.grid_sample(img, mesh, mode="bilinear", padding_mode="border", align_corners=False)


def space_to_depth(x, r):   # SpaceToDepth.forward - frvsr_net.py:178-191
    n, c, h, w = x.shape
    return x.contiguous().view(n, c, h // r, r, w // r, r).permute(0, 1, 3, 5, 2, 4).contiguous().view(n, c * r * r, h // r, w // r)


def frvsr_fnet(a, b, sd, p="fnet"):
    """FNet.forward - frvsr_net.py:108-163"""
    x = torch.cat([a, b], dim=1)
    N, C, H, W = x.shape
    pad = None
    if H % 8 != 0 or W % 8 != 0:
        hd = 8 - H % 8 if H % 8 != 0 else 0
        wd = 8 - W % 8 if W % 8 != 0 else 0
        pad = (wd // 2, wd - wd // 2, hd // 2, hd - hd // 2)
        x = F.pad(x, pad, value=float(x.min()))
    lrelu = lambda v: F.leaky_relu(v, 0.2)
    for i in range(6):
        x = lrelu(_conv(x, sd, f"{p}.body.conv{i + 1}_1", padding=1))
        x = lrelu(_conv(x, sd, f"{p}.body.conv{i + 1}_2", padding=1))
        x = F.max_pool2d(x, 2) if i < 3 else F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    x = lrelu(_conv(x, sd, f"{p}.tail.conv1", padding=1))
    x = torch.tanh(_conv(x, sd, f"{p}.tail.conv2", padding=1))
    if pad is not None:
        w0, wn, h0, hn = pad
        x = x[..., h0:x.size(-2) - hn, w0:x.size(-1) - wn]
    return x


def frvsr_srnet(s2d, lr, sd, p="srnet"):
    """SRNet.forward - frvsr_net.py:66-93 (+ _ResBlock :96-105)"""
    x = F.relu(_conv(torch.cat([s2d, lr], dim=1), sd, f"{p}.head.conv", padding=1))
    i = 0
    while f"{p}.body.{i}.body.conv1.weight" in sd:
        t = F.relu(_conv(x, sd, f"{p}.body.{i}.body.conv1", padding=1))
        x = x + _conv(t, sd, f"{p}.body.{i}.body.conv2", padding=1)
        i += 1
    x = F.relu(_deconv(x, sd, f"{p}.tail.deconv1", stride=2, padding=1, output_padding=1))
    x = F.relu(_deconv(x, sd, f"{p}.tail.deconv2", stride=2, padding=1, output_padding=1))
    return _conv(x, sd, f"{p}.tail.conv", padding=1)


def frvsrnet_forward(inputs, sd, upscale):
    """FRVSRNet.forward - frvsr_net.py:40-61: (sr_imgs, lr_imgs)"""
    sr_imgs, lr_imgs = [], []
    n, c, h, w = inputs[0].shape
    lr_last = inputs[0]
    sr_last = torch.zeros(n, c, h * upscale, w * upscale, dtype=inputs[0].dtype)
    for x in inputs:
        lr_flow = frvsr_fnet(lr_last, x, sd)
        sr_flow = F.interpolate(lr_flow, scale_factor=upscale, mode="bilinear", align_corners=True)
        warped = stn_warp(sr_last.detach(), sr_flow[:, 0], sr_flow[:, 1])
        sr_img = frvsr_srnet(space_to_depth(warped, upscale), x, sd)
        sr_imgs.append(sr_img)
        sr_last = sr_img
        lr_imgs.append(stn_warp(lr_last, lr_flow[:, 0], lr_flow[:, 1]))
        lr_last = x
    return sr_imgs, lr_imgs


# ---- TOFlowNet (toflow_net.py:8-138) ----------------------------------------------------------------------------------
def toflow_warp(x, flow):
    """flow_warp(x, flow, 'bilinear', 'zeros') - toflow_net.py:117-138; flow [N, 2, H, W] here (the reference permutes to
    [N, H, W, 2] before the call, :60,85)"""
    B, C, H, W = x.shape
    gy, gx = torch.meshgrid(torch.arange(0, H), torch.arange(0, W), indexing="ij")
    vx = gx.to(x.dtype) + flow[:, 0]
    vy = gy.to(x.dtype) + flow[:, 1]
    grid = torch.stack((2.0 * vx / max(W - 1, 1) - 1.0, 2.0 * vy / max(H - 1, 1) - 1.0), dim=3)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def _bn2d(x, sd, key, training, buffers=None, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d - toflow_net.py:96-106.  training: batch statistics; `buffers` (a dict of running_mean / running_var
    clones) receives the momentum update exactly as the module's forward does"""
    if training:
        rm = buffers[key + ".running_mean"] if buffers is not None else None
        rv = buffers[key + ".running_var"] if buffers is not None else None
        return F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], True, momentum if buffers is not None else 0.0, eps)
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                        False, 0.0, eps)


def spynet_block(x, sd, p, training, buffers=None):   # SpyNet_Block - toflow_net.py:93-112
    for i in range(4):
        x = _conv(x, sd, f"{p}.block.{3 * i}", padding=3)
        x = F.relu(_bn2d(x, sd, f"{p}.block.{3 * i + 1}", training, buffers))
    return _conv(x, sd, f"{p}.block.12", padding=3)


def spynet(ref, nbr, sd, training, buffers=None, p="spy_net"):   # SpyNet.forward - toflow_net.py:72-90
    B, C, H, W = ref.shape
    refs, nbrs = [ref], [nbr]
    for _ in range(3):
        refs.insert(0, F.avg_pool2d(refs[0], kernel_size=2, stride=2, count_include_pad=False))
        nbrs.insert(0, F.avg_pool2d(nbrs[0], kernel_size=2, stride=2, count_include_pad=False))
    flow = torch.zeros(B, 2, H // 16, W // 16, dtype=ref.dtype)
    for i in range(4):
        flow_up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
        flow = flow_up + spynet_block(torch.cat([refs[i], toflow_warp(nbrs[i], flow_up), flow_up], dim=1), sd,
                                      f"{p}.blocks.{i}", training, buffers)
    return flow


def toflownet_forward(inputs, sd, upscale, training=True, buffers=None):
    """TOFlowNet.forward - toflow_net.py:33-65"""
    T = len(inputs)
    ref_idx = T // 2 if T % 2 == 1 else T // 2 - 1
    x = torch.stack([F.interpolate(f, scale_factor=upscale, mode="bicubic", align_corners=False) for f in inputs], dim=1)
    B, T, C, H, W = x.shape
    pad = None
    if H % 16 != 0 or W % 16 != 0:
        hd = 16 - H % 16 if H % 16 != 0 else 0
        wd = 16 - W % 16 if W % 16 != 0 else 0
        pad = (wd // 2, wd - wd // 2, hd // 2, hd - hd // 2)
        x = F.pad(x, pad, value=float(x.min()))
        B, T, C, H, W = x.shape
    x_ref = x[:, ref_idx]
    warped = []
    for i in range(T):
        if i == ref_idx:
            warped.append(x_ref)
        else:
            warped.append(toflow_warp(x[:, i], spynet(x_ref, x[:, i], sd, training, buffers)))
    y = torch.stack(warped, dim=1).view(B, -1, H, W)
    y = F.relu(_conv(y, sd, "out_block.0", padding=4))
    y = F.relu(_conv(y, sd, "out_block.2", padding=4))
    y = F.relu(_conv(y, sd, "out_block.4"))
    out = _conv(y, sd, "out_block.6") + x_ref
    if pad is not None:
        w0, wn, h0, hn = pad
        out = out[..., h0:out.size(-2) - hn, w0:out.size(-1) - wn]
    return out
