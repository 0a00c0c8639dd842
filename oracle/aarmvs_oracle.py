"""CPU oracle for the AA-RMVSNet per-depth-plane cost-volume pipeline.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch fp32 (CPU) restatement of the reference's
algorithm for the hot path, written as stateless functions over a checkpoint ``state_dict``.  It is
the checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import it.  The product (``aarmvsnet_b200/``) never does.

Parity pin: the reference has no tests or golden vectors of its own, so this
restatement is pinned against outputs of the reference itself run in the build container
(``oracle/make_golden.py`` imports /root/reference unmodified and writes ``tests/golden/*.pt``);
``tests/test_oracle_golden.py`` checks every function here against those fixtures.

Every function cites the reference lines (relative to /root/reference) it restates.
"""
import torch
import torch.nn.functional as F

GN_EPS = 1e-5


def _groups(cout):
    # models/module.py:101  int(max(1, out_channels / group_channel)), group_channel = 8
    return int(max(1, cout / 8))


def conv_gn_relu(x, sd, prefix, stride=1):
    """models/module.py:98-103 (convgnrelu): Conv2d(pad=(k-1)//2) -> GroupNorm -> ReLU."""
    w, b = sd[prefix + ".0.weight"], sd[prefix + ".0.bias"]
    y = F.conv2d(x, w, b, stride=stride, padding=(w.shape[-1] - 1) // 2)
    y = F.group_norm(y, _groups(w.shape[0]), sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], GN_EPS)
    return F.relu(y)


def deform_conv(x, sd, prefix):
    """models/module.py:133-189 (DeformConv2d.forward, modulated, k=3, pad=1, stride=1).

    Offsets: channels 0..8 are ROW offsets, 9..17 COLUMN offsets (module.py:153-159 clamp the first N
    against x.size(2), the last N against x.size(3)); tap k sits at (k // 3 - 1, k % 3 - 1)
    (module.py:191-199, meshgrid 'ij'); sampling happens in the zero-padded image (module.py:142-143)
    whose pixel (i, j) has padded coordinates (i + 1, j + 1) (module.py:201-209).
    """
    b, c, h, w = x.shape
    offset = F.conv2d(x, sd[prefix + ".p_conv.weight"], sd[prefix + ".p_conv.bias"], padding=1)
    mod = torch.sigmoid(F.conv2d(x, sd[prefix + ".m_conv.weight"], sd[prefix + ".m_conv.bias"], padding=1))
    xp = F.pad(x, (1, 1, 1, 1))
    hp, wp = h + 2, w + 2
    k = torch.arange(9)
    tap_r = (k // 3 - 1).view(1, 9, 1, 1).to(x.dtype)
    tap_c = (k % 3 - 1).view(1, 9, 1, 1).to(x.dtype)
    base_r = torch.arange(1, h + 1).view(1, 1, h, 1).to(x.dtype)
    base_c = torch.arange(1, w + 1).view(1, 1, 1, w).to(x.dtype)
    pr = base_r + tap_r + offset[:, :9]      # module.py:218  p = p_0 + p_n + offset
    pc = base_c + tap_c + offset[:, 9:]
    r0 = pr.floor()                          # module.py:150-151
    c0 = pc.floor()
    r1, c1 = r0 + 1, c0 + 1
    r0, r1 = r0.clamp(0, hp - 1), r1.clamp(0, hp - 1)   # module.py:153-154
    c0, c1 = c0.clamp(0, wp - 1), c1.clamp(0, wp - 1)
    pr, pc = pr.clamp(0, hp - 1), pc.clamp(0, wp - 1)   # module.py:159
    # module.py:162-165 bilinear kernel from the CLAMPED corners and the clamped position
    g_lt = (1 + (r0 - pr)) * (1 + (c0 - pc))
    g_rb = (1 - (r1 - pr)) * (1 - (c1 - pc))
    g_lb = (1 + (r0 - pr)) * (1 - (c1 - pc))
    g_rt = (1 - (r1 - pr)) * (1 + (c0 - pc))
    flat = xp.reshape(b, c, hp * wp)

    def take(r, cc):                         # module.py:221-235 gather by r * padded_w + c
        idx = (r.long() * wp + cc.long()).reshape(b, 1, -1).expand(-1, c, -1)
        return flat.gather(2, idx).reshape(b, c, 9, h, w)

    samp = (g_lt.unsqueeze(1) * take(r0, c0) + g_rb.unsqueeze(1) * take(r1, c1)
            + g_lb.unsqueeze(1) * take(r0, c1) + g_rt.unsqueeze(1) * take(r1, c0))   # module.py:174-177
    samp = samp * mod.unsqueeze(1)           # module.py:180-184
    # module.py:186-187,237-243: tap k lands on kernel position (k // 3, k % 3) of a k3/s3 conv
    wgt = sd[prefix + ".conv.weight"].reshape(-1, c * 9)
    out = torch.einsum("ok,bkhw->bohw", wgt, samp.reshape(b, c * 9, h, w))
    return out + sd[prefix + ".conv.bias"].view(1, -1, 1, 1)


def deform_conv_gn_relu(x, sd, prefix):
    """models/module.py:245-250."""
    y = deform_conv(x, sd, prefix + ".0")
    y = F.group_norm(y, _groups(y.shape[1]), sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], GN_EPS)
    return F.relu(y)


def feat_net(img, sd, prefix="feature"):
    """models/drmvsnet.py:41-63 (FeatNet.forward) + :7-24 (IntraViewAAModule.forward)."""
    x = conv_gn_relu(img, sd, prefix + ".init_conv.0")
    x = conv_gn_relu(x, sd, prefix + ".init_conv.1")
    x0 = conv_gn_relu(x, sd, prefix + ".conv0")
    x1 = conv_gn_relu(x0, sd, prefix + ".conv1", stride=2)
    x2 = conv_gn_relu(x1, sd, prefix + ".conv2", stride=2)
    aa = prefix + ".intraAA"
    m0 = conv_gn_relu(deform_conv_gn_relu(x0, sd, aa + ".deformconv0"), sd, aa + ".conv0")
    m1 = conv_gn_relu(deform_conv_gn_relu(x1, sd, aa + ".deformconv1"), sd, aa + ".conv1")
    m2 = conv_gn_relu(deform_conv_gn_relu(x2, sd, aa + ".deformconv2"), sd, aa + ".conv2")
    m1 = F.interpolate(m1, scale_factor=2, mode="bilinear", align_corners=True)
    m2 = F.interpolate(m2, scale_factor=4, mode="bilinear", align_corners=True)
    return torch.cat([m0, m1, m2], 1)


def relative_projection(src_proj, ref_proj):
    """models/module.py:16-18: proj = src_proj @ inverse(ref_proj); rot = proj[:3,:3]; trans = proj[:3,3]."""
    proj = torch.matmul(src_proj, torch.inverse(ref_proj))
    return proj[:, :3, :3], proj[:, :3, 3:4]


def homo_warp(src_fea, src_proj, ref_proj, depth_value):
    """models/module.py:6-38 (homo_warping_depthwise).  depth_value: [B].

    grid_sample is called without align_corners (module.py:36-37): under torch >= 1.3 that means
    align_corners=False although the grid is normalised with the (W-1)/2 formula."""
    b, c, h, w = src_fea.shape
    rot, trans = relative_projection(src_proj, ref_proj)
    dt = src_fea.dtype                 # fp32 like the reference; fp64 only for the tests' noise-floor measurements
    ys, xs = torch.meshgrid(torch.arange(h, dtype=dt), torch.arange(w, dtype=dt), indexing="ij")
    xyz = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, dtype=dt))).unsqueeze(0).repeat(b, 1, 1)
    p = torch.matmul(rot, xyz) * depth_value.view(b, 1, 1) + trans.view(b, 3, 1)
    z = p[:, 2:3, :]
    z = torch.where(z == 0, z + 0.0001, z)                        # module.py:29
    xy = p[:, :2, :] / z
    gx = xy[:, 0, :] / ((w - 1) / 2) - 1                           # module.py:31-32
    gy = xy[:, 1, :] / ((h - 1) / 2) - 1
    grid = torch.stack((gx, gy), dim=2).view(b, h, w, 2)
    return F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def omega(cost_v, sd, prefix="omega.reweight_network"):
    """models/drmvsnet.py:27-38 (InterViewAAModule) with module.py:252-264 (ResnetBlockGn, k=1)."""
    t = conv_gn_relu(cost_v, sd, prefix + ".0")
    s = prefix + ".1.stem"
    u = conv_gn_relu(t, sd, s + ".0")
    v = F.conv2d(u, sd[s + ".1.weight"], sd[s + ".1.bias"])
    v = F.group_norm(v, _groups(v.shape[1]), sd[s + ".2.weight"], sd[s + ".2.bias"], GN_EPS)
    r = F.relu(v + t)
    return torch.sigmoid(F.conv2d(r, sd[prefix + ".2.weight"], sd[prefix + ".2.bias"]))


def aggregate_plane(ref_fea, src_feas, ref_proj, src_projs, depth_value, sd):
    """models/drmvsnet.py:307-318: sum_v (1 + omega(c_v)) * c_v / S with c_v = (warp_v - ref)^2."""
    acc = None
    for src_fea, src_proj in zip(src_feas, src_projs):
        c = (homo_warp(src_fea, src_proj, ref_proj, depth_value) - ref_fea).pow(2)
        term = (omega(c, sd) + 1) * c
        acc = term if acc is None else acc + term
    return acc / len(src_feas)


def convlstm_cell(x, h, c, sd, prefix):
    """models/module.py:76-92: gates = conv3x3(cat[x, h]); split order i, f, o, g."""
    ch = h.shape[1]
    cc = F.conv2d(torch.cat([x, h], 1), sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], padding=1)
    i, f, o, g = torch.split(cc, ch, dim=1)
    c_next = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_next = torch.sigmoid(o) * torch.tanh(c_next)
    return h_next, c_next


def deconv_gn_relu(x, sd, prefix):
    """models/module.py:269-287: ConvTranspose2d(k3, s2, p1, op1) -> GroupNorm(C/8 groups) -> ReLU."""
    y = F.conv_transpose2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=2, padding=1,
                           output_padding=1)
    return F.relu(F.group_norm(y, _groups(y.shape[1]), sd[prefix + ".gn.weight"], sd[prefix + ".gn.bias"], GN_EPS))


HIDDEN_DIMS = (16, 16, 16, 16, 8)          # models/drmvsnet.py:242
STATE_SCALES = (1, 2, 4, 2, 1)             # models/drmvsnet.py:91


def init_state(batch, height, width, dtype=torch.float32):
    """models/drmvsnet.py:202-206 + module.py:94-96 (zeros; the reference's hard .cuda() dropped)."""
    return [(torch.zeros(batch, ch, height // s, width // s, dtype=dtype), torch.zeros(batch, ch, height // s, width // s, dtype=dtype))
            for ch, s in zip(HIDDEN_DIMS, STATE_SCALES)]


def unet_step(x, state, sd, prefix="cost_regularization"):
    """models/drmvsnet.py:145-167 (UNetConvLSTM.forward, process_sq=True). Returns (cost, new state)."""
    cl = prefix + ".cell_list."
    h0, c0 = convlstm_cell(x, *state[0], sd, cl + "0")
    h1, c1 = convlstm_cell(F.max_pool2d(h0, 2, 2), *state[1], sd, cl + "1")
    h2, c2 = convlstm_cell(F.max_pool2d(h1, 2, 2), *state[2], sd, cl + "2")
    up0 = deconv_gn_relu(h2, sd, prefix + ".deconv_0")
    h3, c3 = convlstm_cell(torch.cat([up0, h1], 1), *state[3], sd, cl + "3")
    up1 = deconv_gn_relu(h3, sd, prefix + ".deconv_1")
    h4, c4 = convlstm_cell(torch.cat([up1, h0], 1), *state[4], sd, cl + "4")
    cost = F.conv2d(h4, sd[prefix + ".conv_0.weight"], sd[prefix + ".conv_0.bias"], padding=1)
    return cost, [(h0, c0), (h1, c1), (h2, c2), (h3, c3), (h4, c4)]


def forward_test(sd, imgs, proj_matrices, depth_values, return_index=False, return_cost=False):
    """models/drmvsnet.py:255-270,299-339 (EMVSNet.forward, test phase, evidential head excluded).

    Returns {"depth", "photometric_confidence"} (+ "index": winning plane per pixel, + "cost_reg": [B,D,H,W])."""
    b, n = imgs.shape[:2]
    feats = [feat_net(imgs[:, v], sd) for v in range(n)]
    h, w = feats[0].shape[2:]
    depth_img = torch.zeros(b, h, w)
    max_prob = torch.zeros(b, h, w)
    exp_sum = torch.zeros(b, h, w)
    index = torch.zeros(b, h, w, dtype=torch.int64)
    state = init_state(b, h, w)
    costs = []
    for d in range(depth_values.shape[1]):
        var = aggregate_plane(feats[0], feats[1:], proj_matrices[:, 0],
                              [proj_matrices[:, v] for v in range(1, n)], depth_values[:, d], sd)
        cost, state = unet_step(-1 * var, state, sd)
        if return_cost:
            costs.append(cost.squeeze(1))
        prob = torch.exp(cost.squeeze(1))                          # drmvsnet.py:324
        upd = max_prob < prob                                       # strict '<': first plane wins ties
        max_prob = torch.where(upd, prob, max_prob)
        depth_img = torch.where(upd, depth_values[:, d].view(b, 1, 1).expand(b, h, w), depth_img)
        index = torch.where(upd, torch.full_like(index, d), index)
        exp_sum = exp_sum + prob
    out = {"depth": depth_img, "photometric_confidence": max_prob / exp_sum}
    if return_index:
        out["index"] = index
    if return_cost:
        out["cost_reg"] = torch.stack(costs, 1)
    return out


def forward_train(sd, imgs, proj_matrices, depth_values):
    """models/drmvsnet.py:272-291 (training phase): softmax over the D regularised cost maps."""
    b, n = imgs.shape[:2]
    feats = [feat_net(imgs[:, v], sd) for v in range(n)]
    h, w = feats[0].shape[2:]
    state = init_state(b, h, w, feats[0].dtype)
    costs = []
    for d in range(depth_values.shape[1]):
        var = aggregate_plane(feats[0], feats[1:], proj_matrices[:, 0],
                              [proj_matrices[:, v] for v in range(1, n)], depth_values[:, d], sd)
        cost, state = unet_step(-1 * var, state, sd)
        costs.append(cost)
    return F.softmax(torch.stack(costs, 1).squeeze(2), dim=1)


def cls_loss(prob_volume, depth_gt, mask, depth_value):
    """models/drmvsnet.py:347-381 (mvsnet_cls_loss): masked one-hot cross entropy + WTA depth."""
    b, d, h, w = prob_volume.shape
    dv = depth_value.view(b, d, 1, 1)
    gt_index = torch.argmin(torch.abs(dv - depth_gt.unsqueeze(1)), dim=1)
    gt_index = torch.round(mask * gt_index.float()).long().unsqueeze(1)
    ce = -torch.log(prob_volume.gather(1, gt_index)).squeeze(1)
    valid = mask.sum(dim=[1, 2]) + 1e-6
    loss = torch.mean((mask * ce).sum(dim=[1, 2]) / valid)
    wta = dv.expand(b, d, h, w).gather(1, prob_volume.argmax(1, keepdim=True)).squeeze(1)
    return loss, wta
